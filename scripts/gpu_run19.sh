mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "panel_from_fasta" > gpurun_out/pytest_panel.log 2>&1; tail -5 gpurun_out/pytest_panel.log
nvidia-smi -q | grep -i "persistence" | head -2
timeout 300 python scripts/cli_phases.py 256 5000000 > gpurun_out/cli_phases.log 2>&1; cat gpurun_out/cli_phases.log | tail -80
