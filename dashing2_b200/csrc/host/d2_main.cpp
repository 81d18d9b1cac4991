// d2_main.cpp -- `dashing2-gpu`: drop-in front-end for `dashing2 sketch` / `dashing2 cmp` on top of libd2gpu.
//
// Keeps the reference's command line (option names of /root/reference/src/options.h:63-171, short options
// "m:p:k:w:c:f:S:F:Q:o:L:CNs2BPWh?ZJGHv" src/sketch_main.cpp:63) and its on-disk formats byte for byte:
//   * stacked sketch file  u64 n | u64 S | f64 card[n] | f64 reg[n][S]  (+ FILE.names.txt)   src/sketch_core.cpp:129-171
//   * per-input cache file  f64 card | f64 reg[S], named by makedest()                        src/fastxmerge.cpp:70-118
//   * --save-kmers FILE.kmer64  u32 alphabet|canon<<8, u32 S, u32 k, u32 w, u64 seed, u64[n][S] src/fastxsketch.cpp:245-265
//   * distance output: raw float32 (--binary-output) or the text table / PHYLIP                src/emitrect.cpp:108-403
// Everything numeric happens in libd2gpu (CUDA); this file only parses arguments, reads FASTA/FASTQ
// records (kseq semantics, bonsai/klib/kseq.h:178), batches them and writes files.  Modes the library
// does not implement fail with its error message -- there is no CPU fallback.
#include <zlib.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <charconv>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>
#include <cerrno>
#include <unistd.h>

#include "../../../include/d2gpu.h"

namespace {

// phase timing on stderr with -v
struct PhaseTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0; int verbosity = 0;
    std::vector<std::pair<std::string, double>> phases;      // for --gpu-stats
    void mark(const char *what) {
        const auto now = std::chrono::steady_clock::now();
        phases.emplace_back(what, std::chrono::duration<double, std::milli>(now - last).count());
        if (verbosity > 0) std::fprintf(stderr, "[dashing2-gpu] %-28s %8.1f ms (total %8.1f ms)\n", what,
                                        std::chrono::duration<double, std::milli>(now - last).count(), std::chrono::duration<double, std::milli>(now - t0).count());
        last = now;
    }
} g_timer;

[[noreturn]] void die(const std::string &m) { std::fprintf(stderr, "dashing2-gpu: %s\n", m.c_str()); std::fflush(stderr); _exit(1); }
// output must not be truncated silently (the reference uses checked_fwrite and throws, src/enums.h): every write, flush and close is checked
void xwrite(const void *p, size_t size, size_t n, std::FILE *fp, const std::string &what) {
    if (n && std::fwrite(p, size, n, fp) != n) die("short write to " + what + ": " + std::strerror(errno));
}
std::FILE *xopen(const std::string &path, const char *mode) {
    std::FILE *fp = std::fopen(path.c_str(), mode);
    if (!fp) die("Failed to open " + path + ": " + std::strerror(errno));
    return fp;
}
void xclose(std::FILE *fp, const std::string &what) {
    if (std::fflush(fp) != 0 || std::ferror(fp)) die("write error on " + what + ": " + std::strerror(errno));
    if (std::fclose(fp) != 0) die("close failed on " + what + ": " + std::strerror(errno));
}
void chk(int rc) { if (rc) die(std::string("libd2gpu: ") + d2g_last_error()); }

struct Opts {
    int k = -1, w = -1, nthreads = 1;
    uint64_t S = 1024, seed = 0;
    bool canon = true, cache = false, save_kmers = false, presketched = false, binary = false, parse_by_seq = false;
    int mode = D2G_MODE_OPMH;            // ONE_PERM default (src/sketch_main.cpp:27)
    int alphabet = 0;                    // 0 = DNA; 20 / 14 / 6 / 8: --protein / --protein14 / --protein6 / --protein8 (src/options.h:328-331)
    int measure = D2G_SIMILARITY;
    int shape = D2G_SYMMETRIC; bool phylip = false;
    int topk = -1;
    bool nn_threshold = false; double min_similarity = 0.;   // --similarity-threshold x / -T x (src/options.h:75,309)
    unsigned count_threshold = 0;          // -m / --count-threshold (src/options.h:83-84,352)
    int nlsh = 2;                          // --nLSH (src/options.h:162-163,380)
    uint64_t cssize = 0;                   // -c / --countsketch-size / --countmin-size (src/options.h:78-79,357)
    int ngpus = 1;                         // --gpus N (not a reference option; also D2G_GPUS): files / output rows sharded over N devices
    double fastcmp = 8.; bool bbit = false;   // --fastcmp/--regsize N, --bbit-sigs (src/options.h:76,101)
    std::string ffile, qfile, outfile, cmpout, outprefix;
    std::string gpu_stats;                 // --gpu-stats FILE (not a reference option): host phases + per-class device kernel times as JSON
    std::string filterset;                 // --filterset PATH[:x] (src/options.h:157,377; src/d2.cpp:45-98)
    std::vector<std::string> paths;
    size_t nq = 0;
    int verbosity = 0;
};

// ---- option parsing (only what feeds the two hot paths; unknown reference options are rejected loudly) ----
Opts parse(int argc, char **argv, bool is_cmp) {
    Opts o;
    auto need = [&](int &i) -> std::string { if (i + 1 >= argc) die(std::string("option ") + argv[i] + " needs an argument"); return argv[++i]; };
    for (int i = 0; i < argc; ++i) {
        std::string a = argv[i];
        std::string val; bool has_eq = false;
        if (a.rfind("--", 0) == 0) { auto e = a.find('='); if (e != std::string::npos) { val = a.substr(e + 1); a = a.substr(0, e); has_eq = true; } }
        auto arg = [&]() { return has_eq ? val : need(i); };
        auto shortarg = [&](const char *s) -> bool { // -k31 or -k 31
            if (a.size() >= 2 && a[0] == '-' && a[1] == s[1] && a[1] != '-') { if (a.size() > 2) { val = a.substr(2); has_eq = true; } return true; } return false; };
        if (a == "--kmer-length" || shortarg("-k")) o.k = std::stoi(arg());
        else if (a == "--window-size" || shortarg("-w")) o.w = std::stoi(arg());
        else if (a == "--sketchsize" || shortarg("-S")) o.S = std::stoull(arg());
        else if (a == "--sketch-size-l2" || shortarg("-L")) o.S = 1ull << std::stoi(arg());
        else if (a == "--threads" || shortarg("-p")) o.nthreads = std::max(1, std::stoi(arg()));
        else if (a == "--ffile" || shortarg("-F")) o.ffile = arg();
        else if (a == "--qfile" || shortarg("-Q")) o.qfile = arg();
        else if (a == "--outfile" || shortarg("-o")) o.outfile = arg();
        else if (a == "--cmpout" || a == "--distout" || a == "--cmp-outfile") o.cmpout = arg();
        else if (a == "--outprefix" || a == "--prefix") o.outprefix = arg();
        else if (a == "--seed") o.seed = std::stoull(arg());
        else if (a == "--filterset") o.filterset = arg();
        else if (a == "--gpu-stats") o.gpu_stats = arg();
        else if (a == "--count-threshold" || a == "--threshold" || shortarg("-m")) o.count_threshold = (unsigned)std::max(0, std::atoi(arg().c_str()));
        else if (a == "--topk" || a == "--top-k" || shortarg("-K")) { o.topk = std::stoi(arg()); o.nn_threshold = false; }
        else if (a == "--similarity-threshold" || shortarg("-T")) { o.min_similarity = std::atof(arg().c_str()); o.nn_threshold = true; o.topk = -1; }
        else if (a == "--fastcmp" || a == "--regsize") {
            o.fastcmp = std::atof(arg().c_str());
            if (o.fastcmp != 8. && o.fastcmp != 4. && o.fastcmp != 2. && o.fastcmp != 1.) die("--fastcmp must have 8, 4, 2, or 1 as the argument. These are the only register sizes supported.");
        }
        else if (a == "--countsketch-size" || a == "--countmin-size" || shortarg("-c")) o.cssize = std::strtoull(arg().c_str(), nullptr, 10);
        else if (a == "--nLSH" || a == "--nlsh") o.nlsh = std::stoi(arg());
        else if (a == "--gpus") o.ngpus = std::max(1, std::stoi(arg()));
        else if (a == "--bbit-sigs") o.bbit = true;
        else if (a == "--binary-output" || a == "--emit-binary" || a == "--binary") o.binary = true;
        else if (a == "--phylip") o.phylip = true;
        else if (a == "--asymmetric-all-pairs" || a == "--asymmetric" || a == "--square") o.shape = D2G_ASYMMETRIC;
        else if (a == "--full-setsketch" || a == "--full") o.mode = D2G_MODE_FULL_SETSKETCH;
        else if (a == "--oneperm-setsketch" || a == "--oneperm" || a == "--one-perm" || a == "--oph" || a == "--doph" || a == "-Z") o.mode = D2G_MODE_OPMH;
        else if (a == "--multiset" || a == "--bagminhash" || a == "--bmh" || a == "--BMH") o.mode = D2G_MODE_BAGMINHASH;
        else if (a == "--prob" || a == "--probs" || a == "--pminhash" || a == "--pmh" || a == "--PMH" || a == "--probminhash" || a == "-P") o.mode = D2G_MODE_PROBMINHASH;
        else if (a == "--no-canon" || a == "-C") o.canon = false;
        else if (a == "--protein" || a == "--protein20" || a == "--enable-protein") { o.alphabet = 20; o.canon = false; }
        else if (a == "--protein14") { o.alphabet = 14; o.canon = false; }
        else if (a == "--protein6") { o.alphabet = 6; o.canon = false; }
        else if (a == "--protein8") { o.alphabet = 8; o.canon = false; }
        else if (a == "--cache" || a == "--cache-sketches" || a == "-W") o.cache = true;
        else if (a == "--save-kmers" || a == "-s") o.save_kmers = true;
        else if (a == "--presketched") o.presketched = true;
        else if (a == "--parse-by-seq") o.parse_by_seq = true;
        else if (a == "--containment") o.measure = D2G_CONTAINMENT;
        else if (a == "--symmetric-containment") o.measure = D2G_SYMMETRIC_CONTAINMENT;
        else if (a == "--mash-distance" || a == "--distance" || a == "--poisson-distance") o.measure = D2G_POISSON_LLR;
        else if (a == "--intersection" || a == "--intersection-size") o.measure = D2G_INTERSECTION;
        else if (a == "--union-size") o.measure = D2G_UNION_SIZE;
        else if (a == "--verbose" || a == "-v") ++o.verbosity;
        else if (a == "-h" || a == "--help" || a == "-?") {
            std::printf("dashing2-gpu %s: drop-in for `dashing2 sketch|cmp` (k<=32 DNA; OPMH / Full SetSketch; dense all-pairs / panel).\n"
                        "Options follow the reference: -k -w -S -p -F -Q -o --cmpout --binary-output --phylip --asymmetric-all-pairs\n"
                        "--full-setsketch --oneperm -C/--no-canon --seed --cache --outprefix --save-kmers --presketched\n"
                        "--containment --symmetric-containment --mash-distance --intersection --union-size --topk --fastcmp --bbit-sigs --parse-by-seq -m/--count-threshold -c/--countsketch-size --gpus\n", d2g_version());
            std::exit(0);
        } else if (!a.empty() && a[0] == '-' && a.size() > 1) die("option " + a + " is not supported by the GPU front-end (see DESIGN.md section 7)");
        else o.paths.push_back(a);
    }
    if (o.k < 0) o.k = o.alphabet == 20 ? 14 : o.alphabet == 14 ? 16 : o.alphabet == 6 ? 24 : o.alphabet == 8 ? 22 : 32;   // nregperitem(rht, 64-bit), src/sketch_main.cpp:70
    if (const char *ev = getenv("D2G_GPUS")) if (o.ngpus == 1) o.ngpus = std::max(1, atoi(ev));
    auto read_list = [](const std::string &f, std::vector<std::string> &dst) {
        std::ifstream ifs(f); if (!ifs) die("No path found at " + f);
        for (std::string l; std::getline(ifs, l);) dst.push_back(l);
    };
    if (!o.ffile.empty()) read_list(o.ffile, o.paths);
    const size_t nref = o.paths.size();
    if (!o.qfile.empty()) read_list(o.qfile, o.paths);
    o.nq = o.paths.size() - nref;
    if (o.nq) o.shape = D2G_PANEL;      // src/options.h: -Q implies PANEL
    if (o.paths.empty()) die("No paths provided. See usage.");
    (void)is_cmp;
    return o;
}

// ---- names: makedest() and friends -------------------------------------------------------------
std::string trim_folder(const std::string &s) { auto p = s.find_last_of('/'); return p == std::string::npos ? s : s.substr(p + 1); }
const char *suffix(int mode) { return mode == D2G_MODE_OPMH ? ".opss" : mode == D2G_MODE_FULL_SETSKETCH ? ".ss" : mode == D2G_MODE_BAGMINHASH ? ".bmh" : ".pmh"; }
std::string makedest(const Opts &o, const std::string &path) {   // src/fastxmerge.cpp:70-118
    std::string ret = path.substr(0, path.find_first_of(' '));
    if (!o.outprefix.empty()) ret = o.outprefix + '/' + trim_folder(path);
    if (o.seed) ret += ".seed" + std::to_string(o.seed);
    if (o.canon) ret += ".rc_canon";
    ret += ".sketchsize" + std::to_string(o.S) + ".k" + std::to_string(o.k);
    if (o.w > o.k) ret += ".w" + std::to_string(o.w);
    if (o.count_threshold > 0) ret += ".ct_threshold" + std::to_string(o.count_threshold);
    const bool counted = o.mode == D2G_MODE_BAGMINHASH || o.mode == D2G_MODE_PROBMINHASH;
    if (counted) ret += o.cssize ? ".CountMinCounting" + std::to_string(o.cssize) : std::string(".ExactCounting");
    ret += '.';
    ret += o.mode == D2G_MODE_BAGMINHASH ? "MultisetSpace" : o.mode == D2G_MODE_PROBMINHASH ? "ProbsetSpace" : "SetSpace";
    // bns::to_string(InputType), bonsai/include/bonsai/rhtraits.h:155-169
    const char *rht = o.alphabet == 20 ? ".PROTEIN20" : o.alphabet == 14 ? ".PROTEIN_14" : o.alphabet == 6 ? ".PROTEIN_6" : o.alphabet == 8 ? ".PROTEIN_3BIT" : ".DNA";
    return ret + rht + suffix(o.mode);
}

// ---- FASTA/FASTQ records, kseq semantics ---------------------------------------------------------
struct FileRecords { std::string seq; std::vector<uint64_t> ends; std::vector<std::string> names; };
void read_fastx(const std::string &path, FileRecords &out, bool want_names = false) {
    gzFile fp = gzopen(path.c_str(), "rb");
    if (!fp) die("Could not open file at " + path + ". Abort!");
    gzbuffer(fp, 1 << 18);
    std::string data; std::vector<char> buf(1 << 20);
    for (int n; (n = gzread(fp, buf.data(), (unsigned)buf.size())) > 0;) data.append(buf.data(), n);
    gzclose(fp);
    const char *p = data.data(), *e = p + data.size();
    auto next_line = [&](const char *&b, const char *&le) { le = (const char *)memchr(b, '\n', e - b); if (!le) le = e; };
    while (p < e) {
        // skip to the next header
        while (p < e && *p != '>' && *p != '@') { const char *le; next_line(p, le); p = le < e ? le + 1 : e; }
        if (p >= e) break;
        const bool fastq = *p == '@';
        const char *le; next_line(p, le);
        if (want_names) {   // kseq: the name runs to the first white space; --parse-by-seq drops a leading '>' of the name itself (src/fastxsketchbyseq.cpp:276-277)
            const char *b = p + 1, *t = b; while (t < le && !isspace((unsigned char)*t)) ++t;
            if (b < t && *b == '>') ++b;
            out.names.emplace_back(b, t - b);
        }
        p = le < e ? le + 1 : e;          // header line
        const size_t start = out.seq.size();
        while (p < e && *p != '>' && *p != '+' && *p != '@') {
            next_line(p, le);
            const char *t = le; while (t > p && (t[-1] == '\r')) --t;
            out.seq.append(p, t - p);
            p = le < e ? le + 1 : e;
        }
        const size_t len = out.seq.size() - start;
        if (fastq && p < e && *p == '+') {
            next_line(p, le); p = le < e ? le + 1 : e;
            size_t got = 0;
            while (p < e && got < len) { next_line(p, le); const char *t = le; while (t > p && t[-1] == '\r') --t; got += t - p; p = le < e ? le + 1 : e; }
        }
        out.ends.push_back(out.seq.size());
    }
}

// ---- float -> text exactly as fmt's "{}" (shortest round trip; exponent form iff exp10 < -4 or >= 16) ----
void fmt_float(float v, std::string &out) {
    if (std::isnan(v)) { out += "nan"; return; }
    if (std::isinf(v)) { out += v < 0 ? "-inf" : "inf"; return; }
    if (v == 0) { out += std::signbit(v) ? "-0" : "0"; return; }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    std::string s(buf, r.ptr);                      // d[.ddd]e[+-]XX
    bool neg = s[0] == '-'; if (neg) s.erase(0, 1);
    const auto epos = s.find('e');
    std::string digits = s.substr(0, epos); const int ex = std::stoi(s.substr(epos + 1));
    digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
    if (neg) out += '-';
    const int nd = (int)digits.size();
    if (ex < -4 || ex >= 16) {
        out += digits[0];
        if (nd > 1) { out += '.'; out.append(digits, 1, std::string::npos); }
        char eb[16]; std::snprintf(eb, sizeof eb, "e%c%02d", ex < 0 ? '-' : '+', std::abs(ex)); out += eb;
    } else if (ex >= 0) {
        if (nd <= ex + 1) { out += digits; out.append(ex + 1 - nd, '0'); }
        else { out.append(digits, 0, ex + 1); out += '.'; out.append(digits, ex + 1, std::string::npos); }
    } else { out += "0."; out.append(-ex - 1, '0'); out += digits; }
}

// The CUDA context comes up on its own thread (1-2 s on a B200 box) while the host threads read and parse the first batch.
struct LazyCtx {
    d2g_ctx *ctx = nullptr; int rc = 0; std::string err; std::thread th; bool joined = false;
    void start(int device = 0) {
        th = std::thread([this, device] {
            rc = d2g_init(&ctx, device);
            // test knob: with D2G_SHARE_DEVICE set, contexts beyond the devices present share device 0 (own stream and scratch each), so the
            // file / row sharding of --gpus N can be exercised on a one-GPU box
            if (rc == D2G_EINVAL && device > 0 && getenv("D2G_SHARE_DEVICE")) rc = d2g_init(&ctx, 0);
            if (rc) err = d2g_last_error();
        });
    }
    d2g_ctx *get() {
        if (!joined) { th.join(); joined = true; g_timer.mark("d2g_init (overlapped)"); if (rc) die("libd2gpu: " + err); }
        return ctx;
    }
};

// One context per device (--gpus N): the sketch phase hands whole batches of files to whichever device is free, the compare phase
// gives every device a range of output rows -- the reference's own decomposition (files: src/fastxsketch.cpp:302, rows:
// src/emitrect.cpp:198-326), without any device-to-device traffic because the front-end keeps the sketches in host memory.
struct Gpus {
    std::vector<std::unique_ptr<LazyCtx>> c;
    void start(int n) { for (int g = 0; g < n; ++g) { c.emplace_back(new LazyCtx); c.back()->start(g); } }
    size_t size() const { return c.size(); }
    d2g_ctx *get(size_t g = 0) { return c[g]->get(); }
    LazyCtx &lazy(size_t g = 0) { return *c[g]; }
};

struct Sketches { std::vector<double> sig, card; std::vector<uint64_t> ids; std::vector<std::string> names; uint64_t S = 0; int mode = D2G_MODE_OPMH; };

d2g_sketch_params sketch_params(const Opts &o) {
    d2g_sketch_params p{};
    p.k = o.k; p.w = o.w; p.canon = o.canon; p.mode = o.mode; p.sketchsize = (uint32_t)o.S; p.count_threshold = o.count_threshold;
    p.alphabet = o.alphabet;
    if (o.mode == D2G_MODE_BAGMINHASH || o.mode == D2G_MODE_PROBMINHASH) p.countsketch_size = o.cssize;   // set sketches never count (src/fastxsketch.cpp:425-427)
    p.xormask = 0;
    if (o.seed) { // Wang(seed), src/enums.cpp:133-140
        uint64_t key = o.seed; key = (~key) + (key << 21); key ^= key >> 24; key = (key + (key << 3)) + (key << 8); key ^= key >> 14;
        key = (key + (key << 2)) + (key << 4); key ^= key >> 28; key += key << 31; p.xormask = key;
    }
    return p;
}

// ---- --filterset PATH[:x] (src/d2.cpp:45-98): hashed k-mers of PATH never reach a sketch ------------------------------------------
// PATH alone (or PATH:K): a FASTX file hashed with the options of the run.  A colon followed by anything else selects the reference's raw
// file of 64-bit hashed values -- which the reference binary cannot open (it passes the empty decompression command to fopen, d2.cpp:60-62,
// and aborts); reproduced as an error here, with the reference's message.
uint64_t g_filterset_n = 0;   // hashed values in the filter set, duplicates included (the reference's data_.size())
void load_filterset(Gpus &gpus, const Opts &o) {
    const size_t colon = o.filterset.find_last_of(':');
    const std::string path = o.filterset.substr(0, colon);
    if (colon != std::string::npos && colon + 1 < o.filterset.size() && (o.filterset[colon + 1] & 0xdf) != 'K')
        die("Failed to open file " + path + " for reading");
    if (o.mode == D2G_MODE_BAGMINHASH || o.mode == D2G_MODE_PROBMINHASH)
        die("--filterset with --multiset / --prob: the reference binary crashes on this combination (v2.1.20); refusing instead");
    FileRecords fr;
    read_fastx(path, fr);
    std::vector<uint64_t> off{0};
    for (uint64_t e : fr.ends) off.push_back(e);
    fr.seq.append(64, '\0');
    d2g_sketch_params p = sketch_params(o);
    for (size_t g = 0; g < gpus.size(); ++g) chk(d2g_set_filterset(gpus.get(g), &p, fr.seq.data(), off.data(), off.size() - 1, &g_filterset_n));
    g_timer.mark("filter set");
}

// ---- sketch all inputs through libd2gpu in batches -----------------------------------------------
void sketch_inputs(Gpus &gpus, const Opts &o, Sketches &sk) {
    const size_t n = o.paths.size(), S = o.S;
    sk.S = S; sk.mode = o.mode; sk.names = o.paths;
    sk.sig.assign(n * S, 0.); sk.card.assign(n, 0.);
    if (o.save_kmers) sk.ids.assign(n * S, 0);
    d2g_sketch_params p = sketch_params(o);
    if (o.alphabet) p.alphabet = 0;   // per FILE nothing is fed to the sketch for protein input (see below): the records go down empty
    std::vector<char> todo(n, 1);
    if (o.cache) {   // cache hit = load the per-file sketch (src/fastxsketch.cpp:327-373)
        for (size_t i = 0; i < n; ++i) {
            const std::string dest = makedest(o, o.paths[i]);
            struct stat st;
            if (::stat(dest.c_str(), &st) == 0 && (size_t)st.st_size == 8 + S * 8 && !o.save_kmers) {
                std::FILE *fp = std::fopen(dest.c_str(), "rb");
                if (fp && std::fread(&sk.card[i], 8, 1, fp) == 1 && std::fread(&sk.sig[i * S], 8, S, fp) == S) todo[i] = 0;
                if (fp) std::fclose(fp);
            }
        }
    }
    const size_t G = gpus.size();
    // several devices: smaller batches so that every device gets work (the reference balances files over threads the same way)
    size_t total_bytes = 0;
    if (G > 1) for (size_t j = 0; j < n; ++j) if (todo[j]) { struct stat st; total_bytes += ::stat(o.paths[j].c_str(), &st) == 0 ? (size_t)st.st_size : 0; }
    const size_t batch_bytes = G > 1 ? std::min<size_t>(size_t(1) << 30, std::max<size_t>(size_t(32) << 20, total_bytes / (2 * G) + 1)) : size_t(1) << 30;
    size_t i = 0;
    std::mutex batch_mu;
    auto next_batch = [&](std::vector<size_t> &idx) {   // a batch of files, in input order
        std::lock_guard<std::mutex> lk(batch_mu);
        idx.clear(); size_t est = 0;
        while (i < n && (idx.empty() || est < batch_bytes)) {
            if (todo[i]) { struct stat st; est += ::stat(o.paths[i].c_str(), &st) == 0 ? (size_t)st.st_size : 0; idx.push_back(i); }
            ++i;
        }
        return !idx.empty();
    };
    const unsigned threads_per_worker = (unsigned)std::max<size_t>(1, o.nthreads / G);
    auto worker = [&](size_t g) {
      std::vector<size_t> idx;
      while (next_batch(idx)) {
        std::vector<FileRecords> recs(idx.size());
        {
            std::vector<std::thread> th; std::atomic<size_t> next{0};
            const unsigned nt = (unsigned)std::min<size_t>(threads_per_worker, idx.size());
            for (unsigned t = 0; t < nt; ++t) th.emplace_back([&] { for (size_t j; (j = next++) < idx.size();) read_fastx(o.paths[idx[j]], recs[j]); });
            for (auto &t : th) t.join();
        }
        // Protein input per FILE: the reference binary feeds nothing to the sketch in this mode (empty registers; pinned by the prot*
        // fixtures of tests/golden/make_golden_protein.py) -- only --parse-by-seq sketches residues.  Reproduced, not improved on.
        if (o.alphabet) for (auto &fr : recs) { fr.seq.clear(); fr.ends.assign(fr.ends.size(), 0); }
        if (G == 1) g_timer.mark("read + parse batch");
        // Sub-batches by the bases actually read (the on-disk size says little for .gz): the counting sketches and -m sort a whole batch at
        // once (at most 2^32 bases, ~36 bytes of scratch per base), everything else is only bounded by device memory.
        const bool whole_batch_sort = o.mode == D2G_MODE_BAGMINHASH || o.mode == D2G_MODE_PROBMINHASH || (o.mode == D2G_MODE_OPMH && o.count_threshold > 1);
        const bool element_stream = o.k > 32 || (o.w > o.k && !o.canon);   // stream_kernels.cuh: 8-16 bytes of elements per base on the device
        const uint64_t max_bases = (whole_batch_sort || element_stream) ? 2000000000ULL : 16000000000ULL;
        for (size_t g0 = 0; g0 < idx.size();) {
            size_t g1 = g0; uint64_t tot = 0;
            while (g1 < idx.size() && (g1 == g0 || tot + recs[g1].seq.size() <= max_bases)) tot += recs[g1++].seq.size();
            const size_t nf = g1 - g0;
            // the batch goes to the device packed (2 bits + 1 invalid bit per base), packed by the library's host threads from the per-file
            // buffers: no concatenated ASCII copy, a quarter of the bytes over PCIe
            std::vector<const char *> pieces(nf); std::vector<uint64_t> plen(nf);
            std::vector<uint64_t> off{0}; std::vector<uint32_t> ent;
            uint64_t at = 0;
            for (size_t j = 0; j < nf; ++j) {
                const FileRecords &fr = recs[g0 + j];
                pieces[j] = fr.seq.data(); plen[j] = fr.seq.size();
                for (uint64_t e : fr.ends) { off.push_back(at + e); ent.push_back((uint32_t)j); }
                at += fr.seq.size();
            }
            const uint64_t nw = d2g_packed_words(at);
            std::unique_ptr<uint64_t[]> codes(new uint64_t[nw]); std::unique_ptr<uint32_t[]> mask(new uint32_t[nw]);
            uint64_t nzw = 0;
            chk(d2g_pack_sequences(pieces.data(), plen.data(), nf, codes.get(), mask.get(), &nzw));
            for (size_t j = 0; j < nf; ++j) recs[g0 + j] = FileRecords();
            if (G == 1) g_timer.mark("pack batch");
            const uint32_t ne = (uint32_t)nf;
            std::vector<double> sig((size_t)ne * S), card(ne); std::vector<uint64_t> ids(o.save_kmers ? (size_t)ne * S : 0);
            chk(d2g_sketch_batch_packed(gpus.get(g), &p, codes.get(), nzw ? mask.get() : nullptr, off.data(), ent.data(), ent.size(), ne, nullptr,
                                        sig.data(), card.data(), o.save_kmers ? ids.data() : nullptr, nullptr));
            if (G == 1) g_timer.mark("d2g_sketch_batch_packed");
            for (size_t jj = 0; jj < nf; ++jj) {
                const size_t j = g0 + jj;
                std::copy(sig.begin() + jj * S, sig.begin() + (jj + 1) * S, sk.sig.begin() + idx[j] * S);
                sk.card[idx[j]] = card[jj];
                if (o.save_kmers) std::copy(ids.begin() + jj * S, ids.begin() + (jj + 1) * S, sk.ids.begin() + idx[j] * S);
                if (o.cache) {
                    const std::string dest = makedest(o, o.paths[idx[j]]);
                    std::FILE *fp = xopen(dest, "wb");
                    xwrite(&card[jj], 8, 1, fp, dest); xwrite(&sig[jj * S], 8, S, fp, dest); xclose(fp, dest);
                }
            }
            g0 = g1;
        }
      }
    };
    if (G == 1) worker(0);
    else {
        std::vector<std::thread> ws;
        for (size_t g = 0; g < G; ++g) ws.emplace_back(worker, g);
        for (auto &t : ws) t.join();
        g_timer.mark("sketch batches (all devices)");
    }
}

// ---- --parse-by-seq: one sketch per record of ONE file (fastx2sketch_byseq, src/fastxsketchbyseq.cpp:102-283,284-531) ----
// Records are the entities of d2g_sketch_batch; set sketches whose cardinality estimate is below 10 * S get the exact number
// of distinct k-mers instead (:405-430), from d2g_distinct_kmers.
void sketch_by_seq(LazyCtx &lctx, const Opts &o, Sketches &sk) {
    if (o.paths.size() != 1) die("parse-by-seq currently only handles one file at a time. To process multiple files, simply concatenate them into one file, and run dashing2 on that.");
    const size_t S = o.S;
    FileRecords recs;
    read_fastx(o.paths[0], recs, true);
    g_timer.mark("read + parse records");
    const size_t n = recs.ends.size();
    sk.S = S; sk.mode = o.mode; sk.names = std::move(recs.names);
    sk.sig.assign(n * S, 0.); sk.card.assign(n, 0.);
    if (o.save_kmers) sk.ids.assign(n * S, 0);
    d2g_sketch_params p = sketch_params(o);
    const bool set_space = o.mode == D2G_MODE_OPMH || o.mode == D2G_MODE_FULL_SETSKETCH;
    // The reference's per-thread copies of the sketcher are made by OptSketcher's copy constructor, which builds a fresh
    // OPSetSketch(size) (src/fastxsketchbyseq.cpp:36,46,220-225): the one-permutation sketch loses its mincount, so -m has no
    // effect on set sketches here (verified against the binary, tests/golden/make_golden_mincount.py).  Counting sketches keep it (:462,470).
    if (o.mode == D2G_MODE_OPMH) p.count_threshold = 0;
    const size_t max_bases = size_t(1) << 30, max_ent = std::max<size_t>(1, (size_t(1) << 31) / (S * 8));
    recs.seq.append(64, '\0');
    std::vector<uint64_t> off; std::vector<uint32_t> ent;
    for (size_t r0 = 0; r0 < n;) {
        const uint64_t b0 = r0 ? recs.ends[r0 - 1] : 0;
        size_t r1 = r0 + 1;
        while (r1 < n && r1 - r0 < max_ent && recs.ends[r1] - b0 <= max_bases) ++r1;
        const uint32_t ne = (uint32_t)(r1 - r0);
        off.assign(1, 0); ent.clear();
        for (size_t r = r0; r < r1; ++r) { off.push_back(recs.ends[r] - b0); ent.push_back((uint32_t)(r - r0)); }
        chk(d2g_sketch_batch(lctx.get(), &p, recs.seq.data() + b0, off.data(), ent.data(), ne, ne, nullptr, sk.sig.data() + r0 * S, sk.card.data() + r0,
                             o.save_kmers ? sk.ids.data() + r0 * S : nullptr, nullptr));
        if (set_space) {
            std::vector<size_t> small;
            for (size_t r = r0; r < r1; ++r) {
                if (std::isnan(sk.card[r])) sk.card[r] = 0.;
                if (sk.card[r] < 10. * S) small.push_back(r);
            }
            if (!small.empty()) {
                std::string sub; std::vector<uint64_t> soff{0}; std::vector<uint32_t> sent;
                for (size_t r : small) {
                    const uint64_t b = r ? recs.ends[r - 1] : 0;
                    sub.append(recs.seq, b, recs.ends[r] - b);
                    soff.push_back(sub.size()); sent.push_back((uint32_t)sent.size());
                }
                sub.append(64, '\0');
                std::vector<uint64_t> distinct(small.size());
                chk(d2g_distinct_kmers(lctx.get(), &p, sub.data(), soff.data(), sent.data(), sent.size(), (uint32_t)sent.size(), distinct.data()));
                for (size_t t = 0; t < small.size(); ++t) sk.card[small[t]] = (double)distinct[t];
            }
        }
        r0 = r1;
    }
    g_timer.mark("sketch records");
}

void write_stacked(const Opts &o, const Sketches &sk) {
    const uint64_t n = sk.card.size(), S = sk.S;
    std::FILE *fp = xopen(o.outfile, "wb");
    xwrite(&n, 8, 1, fp, o.outfile); xwrite(&S, 8, 1, fp, o.outfile);
    xwrite(sk.card.data(), 8, n, fp, o.outfile); xwrite(sk.sig.data(), 8, n * S, fp, o.outfile); xclose(fp, o.outfile);
    const std::string np = o.outfile + ".names.txt";
    fp = xopen(np, "wb");
    std::fputs("#Name\tCardinality\n", fp);
    for (size_t i = 0; i < n; ++i) std::fprintf(fp, "%s\t%0.24g\n", sk.names[i].c_str(), sk.card[i]);
    xclose(fp, np);
    if (o.save_kmers) {
        const std::string kp = o.outfile + ".kmer64";
        fp = xopen(kp, "wb");
        // the InputType enumerator (rhtraits.h:7-20): DNA 0, PROTEIN20 2, PROTEIN_3BIT 3, PROTEIN_14 4, PROTEIN_6 5
        const uint32_t rht = o.alphabet == 20 ? 2u : o.alphabet == 8 ? 3u : o.alphabet == 14 ? 4u : o.alphabet == 6 ? 5u : 0u;
        const uint32_t hdr[4] = {rht | (uint32_t(o.canon) << 8), (uint32_t)S, (uint32_t)o.k, (uint32_t)(o.w < 0 ? o.k : o.w)};
        xwrite(hdr, 4, 4, fp, kp); xwrite(&o.seed, 8, 1, fp, kp); xwrite(sk.ids.data(), 8, n * S, fp, kp); xclose(fp, kp);
        fp = xopen(kp + ".names.txt", "wb");
        for (auto &nm : sk.names) { std::fputs(nm.c_str(), fp); std::fputc('\n', fp); }
        xclose(fp, kp + ".names.txt");
    }
}

void load_stacked(const std::string &path, Sketches &sk) {   // src/cmp_main.cpp:24-198 (single stacked file branch)
    std::FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp) die("Failed to open " + path);
    uint64_t n, S;
    if (std::fread(&n, 8, 1, fp) != 1 || std::fread(&S, 8, 1, fp) != 1) die("short stacked sketch file " + path);
    sk.S = S; sk.card.resize(n); sk.sig.resize(n * S);
    if (std::fread(sk.card.data(), 8, n, fp) != n || std::fread(sk.sig.data(), 8, n * S, fp) != n * S) die("truncated stacked sketch file " + path);
    std::fclose(fp);
    std::ifstream ifs(path + ".names.txt");
    for (std::string l; std::getline(ifs, l);) { if (l.empty() || l[0] == '#') continue; sk.names.push_back(l.substr(0, l.find('\t'))); }
    auto ends = [&](const char *s) { const size_t L = strlen(s); return path.size() >= L && path.compare(path.size() - L, L, s) == 0; };
    sk.mode = ends(".opss") ? D2G_MODE_OPMH : ends(".bmh") ? D2G_MODE_BAGMINHASH : ends(".pmh") ? D2G_MODE_PROBMINHASH : D2G_MODE_FULL_SETSKETCH; // cmp_main.cpp:318-351
}

std::string options_string(const Opts &o, int mode) {   // Dashing2Options::to_string, src/d2.cpp:10-43
    std::string r = "Dashing2Options;k:" + std::to_string(o.k);
    if (o.w > 0) r += ";w:" + std::to_string(o.w);
    r += o.parse_by_seq ? ";parsebyseq" : ";parsebyfile"; r += ";trimchr;sketchsize:" + std::to_string(o.S);
    if (o.count_threshold > 0) r += ";" + std::to_string(o.count_threshold);
    r += ";sketchtype:";
    r += mode == D2G_MODE_OPMH ? "onepermsetsketch" : mode == D2G_MODE_FULL_SETSKETCH ? "fullsetsketch" : mode == D2G_MODE_BAGMINHASH ? "bagminhash" : "probminhash";
    r += ";Fastx";
    if (!o.outprefix.empty()) r += ";outprefix:" + o.outprefix;
    if (o.cssize) r += ";counting=countsketch" + std::to_string(o.cssize) + "\n";   // the newline is the reference's (src/d2.cpp:34)
    if (o.canon) r += ";canon";
    if (!o.filterset.empty()) r += ";FilterSetSortedHashSet-size=" + std::to_string(g_filterset_n);   // FilterSet::to_string, src/filterset.h:78-83
    return r;
}

struct Writer {
    const Opts &o; const Sketches &sk; std::FILE *fp; size_t ns, nq; std::string line;
    // several devices: a row range either lands at its byte offset of the binary file (fd >= 0) or is kept in `mem` until the ranges
    // before it have been written
    int fd = -1; uint64_t file_off = 0; std::string *mem = nullptr;
    bool put(const void *p, size_t nbytes) {
        if (mem) { mem->append((const char *)p, nbytes); return true; }
        if (fd >= 0) {   // pwrite may move less than asked (2 GiB per call at most): continue until the block is out
            const char *q = (const char *)p; size_t left = nbytes;
            while (left) {
                const ssize_t w = ::pwrite(fd, q, left, (off_t)file_off);
                if (w < 0 && errno == EINTR) continue;
                if (w <= 0) return false;
                q += w; left -= (size_t)w; file_off += (uint64_t)w;
            }
            return true;
        }
        return std::fwrite(p, 1, nbytes, fp) == nbytes;
    }
    static int sink(void *u, const float *blk, uint64_t first_row, uint64_t n_rows, uint64_t n_vals) {
        Writer *w = (Writer *)u;
        if (w->o.binary) return w->put(blk, 4 * n_vals) ? 0 : 1;
        const float *p = blk;
        for (uint64_t i = first_row; i < first_row + n_rows; ++i) {   // src/emitrect.cpp:172-187
            std::string &l = w->line; l.clear();
            std::string fn = i < w->sk.names.size() && !w->sk.names[i].empty() ? w->sk.names[i] : "E" + std::to_string(i);
            if (fn.size() < 9) fn.append(9 - fn.size(), ' ');
            l += fn;
            const size_t jend = w->o.shape == D2G_PANEL ? w->nq : w->o.shape == D2G_ASYMMETRIC ? w->ns : w->ns - i - 1;
            if (w->o.shape == D2G_SYMMETRIC && !w->o.phylip) for (uint64_t t = 0; t < i + 1; ++t) l += "\t-";
            for (size_t j = 0; j < jend; ++j) { l += '\t'; fmt_float(*p++, l); }
            l += '\n';
            if (!w->put(l.data(), l.size())) return 1;
        }
        return 0;
    }
};

// output rows [b[g], b[g+1]) for device g: equal numbers of pairs (row i of the condensed triangle holds n-1-i of them), else equal rows
std::vector<uint64_t> row_ranges(uint64_t nrows, uint64_t n, int shape, size_t G) {
    std::vector<uint64_t> b(G + 1, nrows);
    b[0] = 0;
    if (shape != D2G_SYMMETRIC) { for (size_t g = 1; g < G; ++g) b[g] = nrows * g / G; return b; }
    const long double total = (long double)n * (n - 1) / 2;
    uint64_t row = 0; long double acc = 0;
    for (size_t g = 1; g < G; ++g) {
        while (row < nrows && acc < total * g / G) { acc += (long double)(n - 1 - row); ++row; }
        b[g] = row;
    }
    return b;
}

void compare_and_emit(Gpus &gpus, const Opts &o, Sketches &sk) {
    d2g_ctx *ctx = gpus.get(0);
    const size_t G = gpus.size();
    const uint64_t n = sk.card.size(), S = sk.S;
    if (sk.mode == D2G_MODE_OPMH) chk(d2g_densify(ctx, sk.sig.data(), sk.ids.empty() ? nullptr : sk.ids.data(), n, (uint32_t)S));  // cmp_core.cpp:686-718
    d2g_cmp_params cp{};
    cp.sketchsize = (uint32_t)S; cp.measure = o.measure; cp.k = o.k; cp.shape = o.shape; cp.n = n; cp.nq = o.nq;
    cp.cmp_kind = (sk.mode == D2G_MODE_OPMH || sk.mode == D2G_MODE_FULL_SETSKETCH) ? D2G_CMP_GTLT : D2G_CMP_EQ;
    // --fastcmp N: make_compressed (src/cmp_core.cpp:741 -> :209-322); the compressed registers replace the signatures
    std::vector<double> creg;
    if (o.fastcmp < 8.) {
        long double a = -1.L, b = -1.L; int32_t used = 0;
        creg.resize(sk.sig.size());
        chk(d2g_make_compressed(sk.sig.data(), sk.ids.size() == sk.sig.size() ? sk.ids.data() : nullptr, n, (uint32_t)S, o.fastcmp, o.bbit ? 1 : 0, &a, &b, creg.data(), &used));
        if (!o.bbit && !used) std::fprintf(stderr, "Truncated via setsketch, a = %0.20Lg and b = %0.24Lg\n", a, b);
        if (!o.bbit && used) std::fprintf(stderr, "Note: setsketch compression yielded infinite value; falling back to b-bit compression\n");
        cp.cmp_kind = used ? D2G_CMP_BBIT : D2G_CMP_SS_COMPRESSED; cp.regbytes = o.fastcmp; cp.compressed_b = b;
    }
    const bool to_stdout = o.cmpout.empty() || o.cmpout[0] == '-';
    std::FILE *fp = to_stdout ? stdout : std::fopen(o.cmpout.c_str(), "wb");
    if (!fp) die("Failed to open path " + o.cmpout + " for writing");
    cp.nlsh = o.nlsh;
    if (o.topk > 0 || o.nn_threshold) {   // KNN / thresholded graph: build_index + refine_results + emit_neighbors (src/cmp_core.cpp:756-799, src/emitnn.cpp:12-52)
        cp.shape = D2G_SYMMETRIC;
        std::vector<uint64_t> indptr(n + 1); uint32_t *idx = nullptr; float *val = nullptr;
        const double *r = sk.sig.data();
        if (cp.cmp_kind == D2G_CMP_EQ && sk.ids.size() == sk.sig.size()) r = reinterpret_cast<const double *>(sk.ids.data());
        // --fastcmp: the index is built over the f64 signatures, refinement compares the compressed registers (src/cmp_core.cpp:741-799)
        const double *ir = creg.empty() ? nullptr : sk.sig.data();
        if (!creg.empty()) r = creg.data();
        const int32_t tk = o.nn_threshold ? -1 : o.topk;
        if (G == 1) chk(d2g_lsh_graph(ctx, &cp, ir, r, sk.card.data(), tk, o.min_similarity, 0, n, indptr.data(), &idx, &val));
        else {   // every device builds the index and scans all queries, but replays / refines / trims only its own range of lists
            std::vector<std::vector<uint64_t>> ip(G); std::vector<uint32_t *> ix(G, nullptr); std::vector<float *> vl(G, nullptr);
            std::vector<std::thread> ws;
            for (size_t g = 0; g < G; ++g) ws.emplace_back([&, g] {
                const uint64_t x0 = n * g / G, x1 = n * (g + 1) / G;
                ip[g].assign(x1 - x0 + 1, 0);
                chk(d2g_lsh_graph(gpus.get(g), &cp, ir, r, sk.card.data(), tk, o.min_similarity, x0, x1, ip[g].data(), &ix[g], &vl[g]));
            });
            for (auto &t : ws) t.join();
            uint64_t tot = 0; for (size_t g = 0; g < G; ++g) tot += ip[g].back();
            idx = (uint32_t *)std::malloc(std::max<uint64_t>(tot, 1) * 4); val = (float *)std::malloc(std::max<uint64_t>(tot, 1) * 4);
            if (!idx || !val) die("out of memory");
            uint64_t at = 0;
            for (size_t g = 0; g < G; ++g) {
                const uint64_t x0 = n * g / G, cnt = ip[g].back();
                for (size_t t = 0; t + 1 < ip[g].size(); ++t) indptr[x0 + t] = at + ip[g][t];
                if (cnt) { std::memcpy(idx + at, ix[g], cnt * 4); std::memcpy(val + at, vl[g], cnt * 4); }
                at += cnt; d2g_free(ix[g]); d2g_free(vl[g]);
            }
            indptr[n] = at;
        }
        const uint64_t nnz = indptr[n];
        if (o.binary) {
            const uint64_t dims[2] = {n, nnz};
            xwrite(dims, 8, 2, fp, o.cmpout); xwrite(indptr.data(), 8, n + 1, fp, o.cmpout); xwrite(idx, 4, nnz, fp, o.cmpout); xwrite(val, 4, nnz, fp, o.cmpout);
        } else {
            std::fputs("#Collection\tNeighbor lists -- name:distance, separated by tabs\n", fp);
            for (uint64_t i = 0; i < n; ++i) {
                std::fputs(i < sk.names.size() ? sk.names[i].c_str() : "", fp);
                for (uint64_t j = indptr[i]; j < indptr[i + 1]; ++j) std::fprintf(fp, "\t%s:%.8g", idx[j] < sk.names.size() ? sk.names[idx[j]].c_str() : "", val[j]);
                std::fputc('\n', fp);
            }
        }
        if (G == 1) { d2g_free(idx); d2g_free(val); } else { std::free(idx); std::free(val); }
        if (!to_stdout) xclose(fp, o.cmpout); else if (std::fflush(fp) != 0) die("write error on standard output");
        return;
    }
    Writer w{o, sk, fp, (size_t)n, (size_t)o.nq, {}};
    if (!o.binary) {   // header, src/emitrect.cpp:136-151
        if (!o.phylip) {
            const char *label = o.shape == D2G_ASYMMETRIC ? "Asymmetric pairwise" : o.shape == D2G_PANEL ? "Panel (Query/Refernce)" : "Symmetric pairwise";
            std::fprintf(fp, "#Dashing2 %s Output\n#Dashing2Options: %s\n#Sources", label, options_string(o, sk.mode).c_str());
            for (uint64_t i = 0; i < n; ++i) { if (i < sk.names.size()) std::fprintf(fp, "\t%s", sk.names[i].c_str()); else std::fprintf(fp, "\tE%llu", (unsigned long long)i); }
            std::fputc('\n', fp);
        } else std::fprintf(fp, "%llu\n", (unsigned long long)n);
    }
    const uint64_t nrows = o.shape == D2G_PANEL ? n - o.nq : n;
    const double *regs = sk.sig.data();
    // equality branch compares the sampled k-mers when they were saved (cmp_core.cpp:501-504)
    if (cp.cmp_kind == D2G_CMP_EQ && sk.ids.size() == sk.sig.size()) regs = reinterpret_cast<const double *>(sk.ids.data());
    if (!creg.empty()) regs = creg.data();
    if (G == 1) chk(d2g_cmp_stream(ctx, &cp, regs, sk.card.data(), 0, nrows, Writer::sink, &w));
    else {
        const std::vector<uint64_t> b = row_ranges(nrows, n, cp.shape, G);
        const bool positioned = o.binary && !to_stdout;      // raw float32 file: every range is written at its own offset
        std::fflush(fp);
        std::vector<std::string> mem(G);
        // Sharded: device g uploads only its block of n / G sketches; the devices exchange register positions and 32-bit ranks over the
        // communicator the library owns (d2g_cmp_stream_sharded).  Falls back to every device holding all registers when no communicator
        // can be had (contexts sharing a device, no NCCL) or the library refuses (NaN registers, S > 65535).
        std::vector<d2g_ctx *> cs(G);
        for (size_t g = 0; g < G; ++g) cs[g] = gpus.get(g);
        bool sharded = S <= 65535 && !getenv("D2G_NO_SHARDED_CMP") && d2g_comm_init_all(cs.data(), (int)G) == D2G_OK;
        for (int attempt = 0; attempt < 2; ++attempt) {
            std::vector<int> rcs(G, 0); std::vector<std::string> errs(G);
            std::vector<std::thread> ws;
            for (size_t g = 0; g < G; ++g) ws.emplace_back([&, g] {
                Writer wg{o, sk, fp, (size_t)n, (size_t)o.nq, {}};
                if (positioned) { uint64_t before = 0; chk(d2g_cmp_rows_size(&cp, 0, b[g], &before)); wg.fd = fileno(fp); wg.file_off = before * 4; }
                else { mem[g].clear(); wg.mem = &mem[g]; }
                if (sharded) {
                    const uint64_t n_per = (n + G - 1) / G, lb = std::min<uint64_t>(n, g * n_per), ln = std::min<uint64_t>(n, (g + 1) * n_per) - lb;
                    rcs[g] = d2g_cmp_stream_sharded(cs[g], &cp, regs + lb * S, sk.card.data() + lb, lb, ln, b[g], b[g + 1], Writer::sink, &wg);
                } else if (b[g + 1] > b[g]) rcs[g] = d2g_cmp_stream(cs[g], &cp, regs, sk.card.data(), b[g], b[g + 1], Writer::sink, &wg);
                if (rcs[g]) errs[g] = d2g_last_error();
            });
            for (auto &t : ws) t.join();
            bool failed = false, unsupported = true;
            for (size_t g = 0; g < G; ++g) if (rcs[g]) { failed = true; unsupported = unsupported && rcs[g] == D2G_EUNSUPPORTED; }
            if (!failed) break;
            if (sharded && unsupported && attempt == 0) { sharded = false; continue; }     // nothing was emitted: the exchange refuses before any row
            for (size_t g = 0; g < G; ++g) if (rcs[g]) die("libd2gpu: " + errs[g]);
        }
        if (o.verbosity > 0) std::fprintf(stderr, "[dashing2-gpu] compare on %zu devices: %s\n", G, sharded ? "sharded registers, NCCL exchange" : "replicated registers");
        if (!positioned) for (size_t g = 0; g < G; ++g) if (std::fwrite(mem[g].data(), 1, mem[g].size(), fp) != mem[g].size()) die("short write to " + o.cmpout);
    }
    if (!to_stdout) xclose(fp, o.cmpout); else if (std::fflush(fp) != 0) die("write error on standard output");
}

} // namespace

// ---- --gpu-stats FILE: what the device did, for the operator (the reference only has -v timers on stderr) ----------------
void write_gpu_stats(Gpus &gpus, const Opts &o) {
    static const char *cls[D2G_T_NCLASSES] = {"sketch_main", "sketch_boot_or_weighted_elements", "compare_tile_or_candidate_scan", "compare_code_preparation",
                                               "pack_ascii", "radix_sorts", "neighbour_refine"};
    std::FILE *fp = xopen(o.gpu_stats, "w");
    std::fprintf(fp, "{\"library\": \"%s\", \"devices\": [", d2g_version());
    for (size_t g = 0; g < gpus.size(); ++g) {
        d2g_ctx *c = gpus.get(g);
        std::fprintf(fp, "%s{\"device\": %zu, \"kernel_launches\": %llu, \"kernel_ms\": {", g ? ", " : "", g, (unsigned long long)d2g_launch_count(c));
        for (int k = 0; k < D2G_T_NCLASSES; ++k) {
            double ms = 0; uint64_t n = 0;
            chk(d2g_get_timing(c, k, &ms, &n));
            std::fprintf(fp, "%s\"%s\": {\"ms\": %.3f, \"timed_regions\": %llu}", k ? ", " : "", cls[k], ms, (unsigned long long)n);
        }
        std::fprintf(fp, "}}");
    }
    std::fprintf(fp, "], \"host_phases_ms\": [");
    for (size_t i = 0; i < g_timer.phases.size(); ++i)
        std::fprintf(fp, "%s{\"phase\": \"%s\", \"ms\": %.3f}", i ? ", " : "", g_timer.phases[i].first.c_str(), g_timer.phases[i].second);
    std::fprintf(fp, "]}\n");
    xclose(fp, o.gpu_stats);
}

// ---- `contain` (src/contain_main.cpp:133-301) --------------------------------------------------------------------------------------
// dashing2 contain [-b] [-p N] [-o OUT] [-F list] DB.kmer64 query...: for every query file, which of each reference's sampled k-mers
// (the FILE.kmer64 of `sketch --save-kmers`) occur in the query's k-mer stream, and how often.  Coverage = sampled k-mers seen / S, mean depth =
// sum of their multiplicities / sampled k-mers seen (uint32 division).  The stream side is d2g_kmer_counts: emit -> sort -> two binary searches
// per sampled k-mer, with the whole database as the ids of one entity.
int contain_main(int argc, char **argv) {
    bool binary = false; std::string outpath; int nthreads = 1;
    for (int i = 0; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&]() -> std::string { if (i + 1 >= argc) die("option " + a + " needs an argument"); return argv[++i]; };
        if (a == "-b") binary = true;
        else if (a == "-h" || a == "-?") { std::printf("dashing2-gpu contain <flags> database.kmer64 <input.fq> <input2.fq>...\n-b: binary output  -o: output path  -F: list of query files  -p: threads\n"); std::exit(0); }
        else if (a.rfind("-p", 0) == 0) nthreads = std::max(1, std::atoi(a.size() > 2 ? a.c_str() + 2 : need().c_str()));
        else if (a.rfind("-o", 0) == 0) outpath = a.size() > 2 ? a.substr(2) : need();
        else if (a.rfind("-F", 0) == 0) { if (a.size() == 2) need(); }
        else if (!a.empty() && a[0] == '-' && a.size() > 1) die("option " + a + " is not a `contain` option (-b -h -p -o -F)");
    }
    (void)nthreads;   // the stream side runs on the device
    std::vector<std::string> pos;
    {   // re-scan: the database is the first non-option argument, -F entries precede the remaining positionals (src/contain_main.cpp:151-153)
        std::vector<std::string> listed, positional;
        for (int i = 0; i < argc; ++i) {
            std::string a = argv[i];
            if (a == "-b") continue;
            if (a == "-p" || a == "-o") { ++i; continue; }
            if (a.rfind("-p", 0) == 0 || a.rfind("-o", 0) == 0) continue;
            if (a == "-F") { std::ifstream ifs(argv[++i]); for (std::string l; std::getline(ifs, l);) listed.push_back(l); continue; }
            if (a.rfind("-F", 0) == 0) { std::ifstream ifs(a.substr(2)); for (std::string l; std::getline(ifs, l);) listed.push_back(l); continue; }
            positional.push_back(a);
        }
        if (positional.empty()) die("usage: dashing2-gpu contain <flags> database.kmer64 <input.fq> ...");
        pos.push_back(positional[0]);
        pos.insert(pos.end(), listed.begin(), listed.end());
        pos.insert(pos.end(), positional.begin() + 1, positional.end());
    }
    const std::string dbpath = pos[0];
    std::vector<std::string> queries(pos.begin() + 1, pos.end());
    if (queries.empty()) die("contain: no query files (reading standard input is not supported by the GPU front-end)");
    // database: u32 alphabet | canon << 8, u32 S, u32 k, u32 w, u64 seed, u64 ids[n][S]
    std::FILE *fp = std::fopen(dbpath.c_str(), "rb");
    if (!fp) die("Failed to open " + dbpath);
    uint32_t hdr[4]; uint64_t seed = 0;
    if (std::fread(hdr, 4, 4, fp) != 4 || std::fread(&seed, 8, 1, fp) != 1) die("short database file " + dbpath);
    std::fseek(fp, 0, SEEK_END); const size_t fsz = (size_t)std::ftell(fp); std::fseek(fp, 24, SEEK_SET);
    const uint32_t S = hdr[1];
    if (!S || (fsz - 24) % S) die("Database corrupted (not a multiple of uint64_t size). Regenerate?");
    std::vector<uint64_t> ids((fsz - 24) / 8);
    if (std::fread(ids.data(), 8, ids.size(), fp) != ids.size()) die("truncated database file " + dbpath);
    std::fclose(fp);
    std::vector<std::string> names;
    {
        std::ifstream ifs(dbpath + ".names.txt");
        if (ifs) for (std::string l; std::getline(ifs, l);) names.push_back(l);
        else { const size_t v = (fsz - 24) / S; while (names.size() < v) names.push_back(std::to_string(v)); }   // the reference's fallback, as it is
    }
    const size_t nitems = names.size();
    if (nitems != ids.size() / S) die("Database corrupted; wrong number of names.");
    if ((hdr[0] & 0xFF) != 0) die("contain: only DNA databases are supported by the GPU front-end");
    if ((uint64_t)nitems * S > 0xFFFFFFFFull) die("contain: database too large for one lookup (n * S must fit 32 bits)");
    Opts o; o.k = (int)hdr[2]; o.w = (int)hdr[3]; o.canon = (hdr[0] & 0x100) != 0; o.seed = seed; o.S = (uint64_t)nitems * S; o.mode = D2G_MODE_OPMH;
    d2g_sketch_params p = sketch_params(o);
    setenv("CUDA_VISIBLE_DEVICES", "0", 0);
    Gpus gpus; gpus.start(1);
    const size_t nq = queries.size();
    std::vector<float> cov(nitems * nq * 2, 0.f);
    float *stats = cov.data() + nitems * nq;
    std::vector<float> counts(ids.size());
    std::vector<uint64_t> total(ids.size());
    for (size_t q = 0; q < nq; ++q) {
        FileRecords fr; read_fastx(queries[q], fr);
        std::fill(total.begin(), total.end(), 0);
        // batches of whole records below the per-call limit; multiplicities add up
        const uint64_t max_bases = 2000000000ULL;
        size_t r0 = 0; const size_t nr = fr.ends.size();
        while (r0 < nr) {
            const uint64_t b0 = r0 ? fr.ends[r0 - 1] : 0;
            size_t r1 = r0 + 1;
            while (r1 < nr && fr.ends[r1] - b0 <= max_bases) ++r1;
            const uint64_t nb = fr.ends[r1 - 1] - b0;
            std::vector<uint64_t> off{0}; std::vector<uint32_t> ent(r1 - r0, 0u);
            for (size_t r = r0; r < r1; ++r) off.push_back(fr.ends[r] - b0);
            const char *piece = fr.seq.data() + b0; const uint64_t plen = nb;
            const uint64_t nw = d2g_packed_words(nb);
            std::unique_ptr<uint64_t[]> codes(new uint64_t[nw]); std::unique_ptr<uint32_t[]> mask(new uint32_t[nw]);
            uint64_t nzw = 0;
            chk(d2g_pack_sequences(&piece, &plen, 1, codes.get(), mask.get(), &nzw));
            chk(d2g_kmer_counts(gpus.get(0), &p, codes.get(), nzw ? mask.get() : nullptr, off.data(), ent.data(), ent.size(), 1, ids.data(), counts.data()));
            for (size_t i = 0; i < ids.size(); ++i) total[i] += (uint64_t)counts[i];
            r0 = r1;
        }
        const double ssiv = 1. / S;
        for (size_t j = 0; j < nitems; ++j) {
            uint32_t matches = 0, sums = 0;
            for (uint32_t t = 0; t < S; ++t) { const uint64_t c = total[j * S + t]; if (c) { ++matches; sums += (uint32_t)c; } }
            if (matches) { cov[nitems * q + j] = (float)(ssiv * matches); stats[nitems * q + j] = (float)(sums / matches); }
        }
    }
    std::FILE *ofp = outpath.empty() ? stdout : xopen(outpath, "w");
    if (binary) {
        const uint64_t dims[2] = {nitems, nq};
        xwrite(dims, 8, 2, ofp, outpath); xwrite(cov.data(), 4, cov.size(), ofp, outpath);
    } else {
        // the reference formats blocks of 16 (AVX-512) / 8 (AVX2) references with fmt::print WITHOUT the file argument, i.e. to standard output
        // whatever -o says (src/contain_main.cpp:262-291); here every entry goes where -o points
        std::fputs("#Dashing2 contain - a list of coverage %%s for the set of references, + mean coverage levels.\n"
                   "#Each matrix entry consists of <coverage%%:mean depth of coverage>\n##References:", ofp);
        for (size_t i = 0; i < nitems; ++i) std::fprintf(ofp, "\t%s", names[i].c_str());
        std::fputc('\n', ofp);
        for (size_t q = 0; q < nq; ++q) {
            std::fputs(queries[q].c_str(), ofp);
            for (size_t j = 0; j < nitems; ++j) std::fprintf(ofp, "\t%.6g%%:%.0f", (double)(100.f * cov[nitems * q + j]), (double)stats[nitems * q + j]);
            std::fputc('\n', ofp);
        }
    }
    if (ofp != stdout) xclose(ofp, outpath);
    gpus.get(0);
    if (std::fflush(nullptr) != 0) die(std::string("flush failed: ") + std::strerror(errno));
    _exit(0);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) die("subcommands: sketch, cmp (alias dist). See `dashing2-gpu sketch -h`");
    const std::string sub = argv[1];
    const bool is_cmp = sub == "cmp" || sub == "dist";
    if (sub == "contain") return contain_main(argc - 2, argv + 2);
    if (!is_cmp && sub != "sketch") die("subcommand '" + sub + "' is outside the GPU hot paths (sketch, cmp and contain are provided)");
    Opts o = parse(argc - 2, argv + 2, is_cmp);
    g_timer.verbosity = o.verbosity;
    // this front-end drives one GPU: hiding the others from the CUDA driver cuts its start-up (context creation touches
    // every visible device) from seconds to a fraction of a second on an 8-GPU box.  An existing setting is respected.
    {
        std::string vis;
        for (int g = 0; g < o.ngpus; ++g) vis += (g ? "," : "") + std::to_string(g);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 0);
    }
    Gpus gpus;
    gpus.start(o.ngpus);
    if (!o.gpu_stats.empty()) for (size_t g = 0; g < gpus.size(); ++g) chk(d2g_set_timing(gpus.get(g), 1));
    Sketches sk;
    if (is_cmp && o.presketched) {
        if (o.paths.size() != 1) die("--presketched: pass one stacked sketch file (the reference's multi-file branch is degenerate for panels, SURVEY 8a b9)");
        load_stacked(o.paths[0], sk);
        o.S = sk.S;
    } else {
        if (!o.filterset.empty()) load_filterset(gpus, o);
        if (o.parse_by_seq) sketch_by_seq(gpus.lazy(0), o, sk); else sketch_inputs(gpus, o, sk);
        if (!o.outfile.empty()) {
            // the reference densifies signatures_ in place before the stacked file is closed when --cmpout is given
            if (!o.cmpout.empty() && sk.mode == D2G_MODE_OPMH) chk(d2g_densify(gpus.get(0), sk.sig.data(), sk.ids.empty() ? nullptr : sk.ids.data(), sk.card.size(), (uint32_t)sk.S));
            write_stacked(o, sk);
        }
    }
    g_timer.mark("sketches ready / written");
    (void)0;
    if (is_cmp || !o.cmpout.empty()) {
        if (is_cmp && o.cmpout.empty()) o.cmpout = "-";
        compare_and_emit(gpus, o, sk);
        g_timer.mark("compare + emit");
    }
    if (!o.gpu_stats.empty()) write_gpu_stats(gpus, o);
    // every output file is closed / flushed; tearing the CUDA context down costs another 0.2-0.5 s and frees nothing the
    // process exit does not free
    for (size_t g = 0; g < gpus.size(); ++g) gpus.get(g);
    if (std::fflush(nullptr) != 0) die(std::string("flush failed: ") + std::strerror(errno));
    _exit(0);
    return 0;
}
