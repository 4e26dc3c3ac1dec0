// buildG -- drop-in replacement for the reference's graph-construction executable (src/BuildGraph/src/main.cpp):
// same options, same config key, same input formats, same output files; the stage itself runs on a B200 through the
// C ABI of include/disco_gpu.h.
//
//   buildG [-pe f1,f2,...] [-se f1,...] -f <prefix> -p <cfg> [-t <threads/shards>] [-m <GB>] [-w <n>] [-g <gpu>[,<gpu>...] | all]
//
// Mirrors: parseArguments (main.cpp:79-150: unknown option -> usage + exit 1; no args / -h -> usage + exit 0),
// readOverlapParameter (main.cpp:152-176: MinOverlap4BuildGraph, default 30, missing file -> exit 1),
// readCheckpointInfo (main.cpp:178-204: GC=Complete -> nothing to do), Dataset's ReadIDMap (Dataset.cpp:103-129),
// the file set runDisco.sh lists for -n <t>.  -t is the number of output shards (and host threads): shard t
// gets the edges of a contiguous range of reads with the reference's mark flags (2 = both endpoints in this shard, 0 / 1 =
// only the source / destination, the edge then also appears in the other endpoint's shard: OverlapGraph.cpp:826-859), so
// that runDisco.sh's one-parsimplify-per-file step runs in parallel; the contained rows go to shard 0 (rows of one
// container must stay together).  -m and -w are accepted and ignored (the reference ignores -w too, OverlapGraph.cpp:59-81).
// New, optional: -g <device>[,<device>...] (or env DISCO_GPUS / DISCO_GPU), default 0.  Several devices = the key-sharded
// partitioning of buildG-MPIRMA inside this one process (disco_gpu_build_graph_multi: reads replicated, table sharded
// by key, adjacency by read range, remote shards read over NVLink).  Fatal errors print the reference's message and, unlike the
// reference (exit(0), Common.h:64), return 1.
#include "../../include/disco_host.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace std;

static double now() { return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count(); }

static vector<string> split_tok(const string &s, char d)
{
    vector<string> out;
    stringstream ss(s);
    string item;
    while (getline(ss, item, d)) out.push_back(item);
    return out;
}
static string trimmed(string s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == string::npos ? string() : s.substr(a, b - a + 1);
}

static void usage()
{
    cerr << endl << "Usage: buildG [OPTION]...[PRARAM]..." << endl;
    cerr << "  -pe\tnumber of files and paired-end file names" << endl;
    cerr << "  -se\tnumber of files and single-end file names" << endl;
    cerr << "  -f\tAll file name prefix" << endl;
    cerr << "  -t\tmaximum threads used (= number of output shards)" << endl;
    cerr << "  -m\tmaximum memory usage allowed (accepted for compatibility)" << endl;
    cerr << "  -g\tCUDA device(s) to run on, comma separated, or 'all' (default 0 or $DISCO_GPUS)" << endl;
}

static future<string> g_ctx_ready; // the GPU contexts, being created on another thread

[[noreturn]] static void die(const string &msg)
{
    if (g_ctx_ready.valid()) g_ctx_ready.wait(); // do not tear the process down under a thread that is inside the CUDA driver
    cout << endl << "Exit from buildG (B200)" << endl << "Message: " << msg << endl;
    exit(1);
}

static void touch(const string &path, const string &content = "")
{
    ofstream f(path.c_str());
    if (!f) die("Unable to open file: " + path);
    f << content;
}

int main(int argc, char **argv)
{
    cout << "Software: Disco Assembler graph construction (B200 / CUDA hot path)" << endl;
    const double t_main = now();
    vector<string> pe, se;
    string prefix, cfg;
    unsigned long long threads = 1;
    string devices = getenv("DISCO_GPUS") ? getenv("DISCO_GPUS") : getenv("DISCO_GPU") ? getenv("DISCO_GPU") : "0";
    cout << "PRINTING ARGUMENTS" << endl;
    for (int i = 0; i < argc; i++) cout << argv[i] << ' ';
    cout << endl;
    if (argc == 1) { usage(); return 0; }
    for (int i = 1; i < argc; i++) {
        const string a = argv[i];
        auto next = [&]() -> string { if (i + 1 >= argc) { usage(); cerr << "Missing value for " << a << endl; exit(1); } return argv[++i]; };
        if (a == "-pe") { for (auto &f : split_tok(next(), ',')) pe.push_back(f); }
        else if (a == "-se") { for (auto &f : split_tok(next(), ',')) se.push_back(f); }
        else if (a == "-f") prefix = next();
        else if (a == "-t") threads = stoull(next(), nullptr, 0);
        else if (a == "-w") (void)next();
        else if (a == "-m") (void)next();
        else if (a == "-p") cfg = next();
        else if (a == "-g") devices = next();
        else {
            usage();
            if (a == "-h" || a == "--help") return 0;
            cerr << "Unknown option: " << a << endl << endl;
            return 1;
        }
    }
    if (threads < 1) threads = 1;

    // MinOverlap4BuildGraph (main.cpp:152-176)
    unsigned long long min_overlap = 30;
    {
        ifstream f(cfg.c_str());
        if (!f.is_open()) { cerr << "Unable to open parameter file: " << cfg << endl; return 1; }
        string line;
        while (getline(f, line)) {
            if (line.find('=') == string::npos) continue;
            vector<string> tok = split_tok(line, '=');
            if (tok.size() >= 2 && trimmed(tok[0]) == "MinOverlap4BuildGraph") min_overlap = stoull(trimmed(tok[1]), nullptr, 0);
        }
    }
    cout << "Minimum overlap length: " << min_overlap << endl;

    // checkpoint (main.cpp:178-204, :48-52)
    {
        ifstream f((prefix + "_CheckpointInfo.txt").c_str());
        string line;
        while (f.is_open() && getline(f, line)) {
            vector<string> tok = split_tok(line, '=');
            if (tok.size() >= 2 && trimmed(tok[0]) == "GC" && trimmed(tok[1]) == "Complete") {
                cout << "Graph already exists. Using previously built graph..." << endl << "Exiting graph construction." << endl;
                return 0;
            }
        }
    }

    // ---- the GPU contexts are created on another thread while the input is parsed (creating a CUDA context takes as long
    // as parsing ten million reads)
    vector<int> devs;
    if (trimmed(devices) == "all") {
        const int nd = disco_gpu_device_count();
        if (nd < 1) die("no CUDA device");
        for (int d = 0; d < nd && d < DISCO_MAX_SHARDS; d++) devs.push_back(d);
    } else {
        for (auto &d : split_tok(devices, ',')) if (!trimmed(d).empty()) devs.push_back(atoi(trimmed(d).c_str()));
    }
    if (devs.empty()) devs.push_back(0);
    if (devs.size() > DISCO_MAX_SHARDS) die("at most " + to_string(DISCO_MAX_SHARDS) + " GPUs");
    vector<disco_ctx *> ctxs(devs.size(), nullptr);
    double t_ctx = 0;
    g_ctx_ready = async(launch::async, [&]() -> string {
        const double t = now();
        for (size_t r = 0; r < devs.size(); r++)
            if (disco_gpu_create(&ctxs[r], devs[r])) return string("GPU context: ") + disco_gpu_last_error(nullptr);
        t_ctx = now() - t;
        return string();
    });

    // ---- Dataset (Dataset.cpp:34-149): paired files first, then single files; ReadIDMap in file-index space
    double t0 = now();
    disco_reads *reads = disco_reads_new((uint32_t)min_overlap, (int)threads);
    {
        ofstream map((prefix + "_ReadIDMap.txt").c_str());
        if (!map) die("Unable to open file: " + prefix + "_ReadIDMap.txt");
        auto load = [&](const vector<string> &files, const char *kind) {
            for (size_t i = 0; i < files.size(); i++) {
                const unsigned long long before = disco_reads_records(reads);
                cout << "Reading dataset from file: " << files[i] << endl;
                if (disco_reads_add_file(reads, files[i].c_str())) die(disco_host_last_error());
                map << files[i] << ": " << kind << " file " << i + 1 << "\nReadID Range: (" << before + 1 << "," << disco_reads_records(reads) << ")\n";
            }
        };
        load(pe, "Paired-end");
        load(se, "Singleton");
    }
    if (disco_reads_finalize(reads)) die(disco_host_last_error());
    const uint64_t n = disco_reads_count(reads);
    cout << setw(10) << n << " good reads in all datasets." << endl;
    cout << setw(10) << disco_reads_records(reads) - n << " bad reads in all datasets." << endl;
    cout << "Shortest read length in all datasets: " << disco_reads_min_len(reads) << endl;
    cout << " Longest read length in all datasets: " << disco_reads_max_len(reads) << endl;
    if (n == 0) die("No reads found in the read files provided! Please check if the filename(s) and path(s) are correct.");
    cout << "Function Dataset() finished in " << now() - t0 << " Seconds." << endl << endl;

    // ---- hot path on the GPU(s)
    t0 = now();
    double t_load = 0, t_graph = 0, t_fetch = 0, t1 = now();
    {
        const string err = g_ctx_ready.get();
        if (!err.empty()) die(err);
    }
    const double t_wait = now() - t1;
    t1 = now();
    for (size_t r = 0; r < devs.size(); r++) {
        if (disco_gpu_load_reads(ctxs[r], disco_reads_packed(reads), disco_reads_len(reads), n, disco_reads_words_per_read(reads)))
            die(string("load reads: ") + disco_gpu_last_error(ctxs[r]));
        t_load += now() - t1; t1 = now();
    }
    if (ctxs.size() == 1) {
        if (disco_gpu_build_graph(ctxs[0], (uint32_t)min_overlap, 4 /* MAX_EDGE_PER_KMER, Common.h:62 */))
            die(string("build graph: ") + disco_gpu_last_error(ctxs[0]));
    } else if (disco_gpu_build_graph_multi(ctxs.data(), (uint32_t)ctxs.size(), (uint32_t)min_overlap, 4)) {
        string msg = "build graph:";
        for (auto c : ctxs) if (*disco_gpu_last_error(c)) msg += string(" [") + disco_gpu_last_error(c) + "]";
        die(msg);
    }
    t_graph = now() - t1; t1 = now();
    // contained rows: every context holds all of them; edges: each context holds those of its read range
    uint64_t n_contained = 0, n_edges = 0, w = 0;
    disco_gpu_counts(ctxs[0], &n_contained, nullptr);
    vector<disco_crow> rows(n_contained);
    if (disco_gpu_get_contained(ctxs[0], rows.data(), rows.size(), &w)) die(disco_gpu_last_error(ctxs[0]));
    vector<uint64_t> first(ctxs.size() + 1, 0);
    for (size_t r = 0; r < ctxs.size(); r++) { uint64_t ne = 0; disco_gpu_counts(ctxs[r], nullptr, &ne); first[r + 1] = first[r] + ne; }
    n_edges = first.back();
    vector<disco_edge> edges(n_edges);
    disco_stats st{};
    for (size_t r = 0; r < ctxs.size(); r++) {
        // sorted by (src, dst) where they lie in HBM; context r holds the edges whose source is in its read range, so the
        // concatenation in context order is sorted too
        if (disco_gpu_sort_edges(ctxs[r])) die(disco_gpu_last_error(ctxs[r]));
        if (disco_gpu_get_edges(ctxs[r], edges.data() + first[r], first[r + 1] - first[r], &w)) die(disco_gpu_last_error(ctxs[r]));
        disco_stats sr;
        disco_gpu_get_stats(ctxs[r], &sr);
        if (r == 0) st = sr;
        else {
            st.raw_directed_edges += sr.raw_directed_edges; st.cap_fired += sr.cap_fired;
            st.multi_overlap_pairs += sr.multi_overlap_pairs; st.one_sided_edges += sr.one_sided_edges;
            st.ms_total = max(st.ms_total, sr.ms_total);
        }
    }
    t_fetch = now() - t1; t1 = now();
    for (auto c : ctxs) disco_gpu_destroy(c);
    const double t_gpu = now() - t0;
    cout << "GPU stage: context " << t_ctx << " s (created while the input was parsed; waited " << t_wait << " s for it), upload " << t_load << " s, graph " << t_graph << " s, sort + download " << t_fetch
         << " s, release " << now() - t1 << " s" << endl;
    cout << "Hash Table size set to: " << st.table_buckets * 4 << endl;
    cout << "Function insertDataset() finished in " << (st.ms_table_all + st.ms_table_nc) / 1000.0 << " Seconds." << endl;
    cout << "Function markContainedReads() finished in " << (st.ms_contained + st.ms_finish_contained) / 1000.0 << " Seconds." << endl;
    cout << endl << setw(10) << n - n_contained << " Non-contained reads. (Keep as is)\n";
    cout << setw(10) << n_contained << " contained reads. (Need to change their mate-pair information)" << endl;
    cout << "GPU: " << st.raw_directed_edges << " directed overlaps found, " << n_edges << " edges after transitive reduction; cap_fired "
         << st.cap_fired << ", multi_overlap_pairs " << st.multi_overlap_pairs << ", one_sided_edges " << st.one_sided_edges << endl;
    cout << "Function buildOverlapGraphFromHashTable() finished in " << t_gpu << " Seconds. (device " << st.ms_total / 1000.0 << ")" << endl << endl;

    // ---- output files
    t0 = now();
    const uint64_t *fi = disco_reads_file_index(reads);
    const uint16_t *len = disco_reads_len(reads);
    disco_host_sort_contained(rows.data(), rows.size(), len, (uint32_t)min_overlap);
    if (disco_write_pargraph_sharded(prefix.c_str(), (uint32_t)threads, edges.data(), edges.size(), n, fi, len)) die(disco_host_last_error());
    for (unsigned long long t = 0; t < threads; t++) {
        const string tp = prefix + "_" + to_string(t);
        if (disco_write_contained((tp + "_containedReads.txt").c_str(), rows.data(), t == 0 ? rows.size() : 0, fi, len, 0)) die(disco_host_last_error());
        touch(tp + "_startRead.txt", t == 0 ? "1\n" : "");
    }
    touch(prefix + "_CheckpointInfo.txt", "CCR=Complete\nGC=Complete\n"); // OverlapGraph.cpp:486-493, main.cpp:64-70
    cout << "Function saveParGraphToFile() finished in " << now() - t0 << " Seconds." << endl;
    cout << endl << "Graph construction complete." << endl;
    cout << "Function main() finished in " << now() - t_main << " Seconds." << endl;
    disco_reads_free(reads);
    return 0;
}
