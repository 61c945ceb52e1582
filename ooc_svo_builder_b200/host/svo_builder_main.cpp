// svo_builder / svo_builder_binary -- host side of the B200-native SVO builder.
//
// Drop-in for the reference command line (src/svo_builder/main.cpp:100-196):
//     svo_builder[_binary] -f <x.tri> [-s g] [-l MB] [-d pct] [-levels]
//                          [-c model|linear|normal|fixed] [-v] [-h]
// Reads <x>.tri / <x>.tridata (libtri formats, tri_tools.h:75-128) and writes
// <x><g>_<P>.octree / .octreenodes / .octreedata with the reference's byte layout
// (octree_io.h:49-83). Two executables from this one source, selected by
// -DBINARY_VOXELIZATION like the reference (CMakeLists.txt:46-49).
//
// All computation happens in libsvo_b200.so (sm_100a kernels) through the C ABI
// in include/svo_b200.h; this file only parses, reads, streams and writes. There
// is no CPU fallback: without a B200 the program reports the library's error.
// Like the reference, every user error prints a message and exits with status 0.
#include "svo_b200.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace {

#ifdef BINARY_VOXELIZATION
constexpr bool kBinary = true;
constexpr int kFloatsPerTri = 9;
#else
constexpr bool kBinary = false;
constexpr int kFloatsPerTri = 21;
#endif

const char* kVersion = "1.6.4-b200";

struct Options {
    std::string tri_path;
    uint64_t gridsize = 1024;        // main.cpp:30
    uint64_t memory_limit = 2048;    // main.cpp:31
    float sparseness = 0.10f;        // main.cpp:32
    int color = SVO_COLOR_MODEL;
    std::string color_name = "Color from model (fallback to fixed color if model has no color)";
    bool levels = false;
    bool verbose = false;
    int device = 0;
    int gpus = 1;                    // -gpus N: shard the partitions over N devices (one context per device, in-process)
    int separability = 26;           // -sep 6: opt-in 6-separating variant (not in the reference)
};

struct TriHeader {
    std::string base;                // path without extension
    int version = 1;
    int geometry_only = 0;
    uint64_t n_triangles = 0;
    float bbox_min[3] = { 0, 0, 0 }, bbox_max[3] = { 0, 0, 0 };
};

struct WallTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

void banner() {
    std::cout << "--------------------------------------------------------------------\n";
    std::cout << "Out-Of-Core SVO Builder " << kVersion << (kBinary ? " - Geometry only version" : "") << "\n";
    std::cout << "B200-native build (sm_100a kernels, " << svo_version() << ")\n";
    std::cout << "--------------------------------------------------------------------\n" << std::endl;
}

void usage() {
    std::cout << "Example: svo_builder -f /home/jeroen/bunny.tri\n\n"
                 "All available program options:\n\n"
                 "-f <filename.tri>     Path to a .tri input file.\n"
                 "-s <gridsize>         Voxel gridsize, should be a power of 2. Default 1024.\n"
                 "-l <memory_limit>     Memory limit for process, in Mb. Default 2048. Decides the partition count.\n"
                 "-levels               Generate intermediary voxel levels by averaging voxel data\n"
                 "-c <option>           Coloring of voxels (Options: model (default), fixed, linear, normal)\n"
                 "-d <percentage>       Percentage of memory limit to be used additionaly for sparseness optimization\n"
                 "-g <device>           CUDA device index (default 0)\n"
                 "-gpus <n>             Shard the build over the first n visible CUDA devices (power of two, default 1)\n"
                 "-sep <26|6>           Separability: 26 = conservative voxelization (default, the reference's), 6 = thin\n"
                 "-v                    Be very verbose.\n"
                 "-h                    Print help and exit." << std::endl;
}

[[noreturn]] void bad_arguments(const char* why = nullptr) {
    if (why) std::cout << why << std::endl;
    std::cout << "Not enough or invalid arguments, please try again.\n"
                 "At the bare minimum, I need a path to a .TRI file\n" << std::endl;
    usage();
    std::exit(0);      // the reference exits 0 on every user error (main.cpp:104-107)
}

bool power_of_two(uint64_t x) { return x != 0 && (x & (x - 1)) == 0; }

Options parse(int argc, char** argv) {
    Options o;
    std::cout << "Reading program parameters ..." << std::endl;
    if (argc < 3) bad_arguments();
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto value = [&]() -> std::string {
            if (i + 1 >= argc) bad_arguments();
            return std::string(argv[++i]);
        };
        if (a == "-f") {
            o.tri_path = value();
            if (o.tri_path.find(".tri") == std::string::npos)
                bad_arguments("Data filename does not end in .tri - I only support that file format");
        } else if (a == "-s") {
            o.gridsize = (uint64_t)std::atoi(value().c_str());
            if (!power_of_two((unsigned)o.gridsize)) bad_arguments("Requested gridsize is not a power of 2");
        } else if (a == "-l") {
            const int v = std::atoi(value().c_str());
            if (v <= 1) bad_arguments("Requested memory limit is nonsensical. Use a value >= 1");   // main.cpp:130-135
            o.memory_limit = (uint64_t)v;
        } else if (a == "-d") {
            o.sparseness = std::atoi(value().c_str()) / 100.0f;      // integer percent, main.cpp:139-140
            if (o.sparseness < 0) bad_arguments("Requested data memory limit is nonsensical. Use a value > 0");
        } else if (a == "-v") {
            o.verbose = true;
        } else if (a == "-levels") {
            o.levels = true;
        } else if (a == "-g") {
            o.device = std::atoi(value().c_str());
        } else if (a == "-sep") {
            o.separability = std::atoi(value().c_str());
            if (o.separability != 6 && o.separability != 26) bad_arguments("Requested separability must be 26 (conservative, default) or 6 (thin)");
        } else if (a == "-gpus") {
            o.gpus = std::atoi(value().c_str());
            if (o.gpus < 1 || (o.gpus & (o.gpus - 1)) != 0 || o.gpus > 16) bad_arguments("Requested GPU count must be a power of 2 (1..16)");
        } else if (a == "-c") {
            const std::string c = value();
            if (kBinary) {
                std::cout << "You asked to generate colors, but we're only doing binary voxelisation." << std::endl;
            } else if (c == "model") {
                o.color = SVO_COLOR_MODEL;
            } else if (c == "linear") {
                o.color = SVO_COLOR_LINEAR; o.color_name = "Linear";
            } else if (c == "normal") {
                o.color = SVO_COLOR_NORMAL; o.color_name = "Normal";
            } else if (c == "fixed") {
                o.color = SVO_COLOR_FIXED; o.color_name = "Fixed";
            } else {
                std::cout << "Unrecognized color switch: " << c << ", so reverting to colors from model." << std::endl;
            }
        } else if (a == "-h") {
            usage();
            std::exit(0);
        } else {
            bad_arguments();
        }
    }
    if (o.tri_path.empty()) bad_arguments();
    if (o.verbose) {
        std::cout << "  filename: " << o.tri_path << "\n  gridsize: " << o.gridsize << "\n  memory limit: " << o.memory_limit
                  << "\n  sparseness optimization limit: " << o.sparseness << " resulting in " << (o.sparseness * o.memory_limit)
                  << " memory limit.\n  color type: " << o.color_name << "\n  generate levels: " << o.levels
                  << "\n  verbosity: " << o.verbose << std::endl;
    }
    return o;
}

bool file_exists(const std::string& p) {
    if (FILE* f = std::fopen(p.c_str(), "rb")) { std::fclose(f); return true; }
    return false;
}

// .tri text header: `#tri <version>` then keyword/value pairs until END (tri_tools.h:75-115)
bool read_tri_header(const std::string& path, TriHeader& h) {
    std::ifstream in(path.c_str());
    if (!in) { std::cout << "  Error: file " << path << " does not exist." << std::endl; return false; }
    h.base = path.substr(0, path.find_last_of('.'));
    std::string word;
    in >> word;
    if (word != "#tri") { std::cout << "  Error: first line reads [" << word << "] instead of [#tri]" << std::endl; return false; }
    in >> h.version;
    bool done = false;
    while (in.good() && !done) {
        in >> word;
        if (word == "END") done = true;
        else if (word == "ntriangles") in >> h.n_triangles;
        else if (word == "geo_only") in >> h.geometry_only;
        else if (word == "bbox") in >> h.bbox_min[0] >> h.bbox_min[1] >> h.bbox_min[2] >> h.bbox_max[0] >> h.bbox_max[1] >> h.bbox_max[2];
        else {
            std::cout << "  unrecognized keyword [" << word << "], skipping" << std::endl;
            in.ignore(1 << 20, '\n');
        }
    }
    if (!done) { std::cout << "  error reading header" << std::endl; return false; }
    return true;
}

[[noreturn]] void die(svo_ctx* ctx, const char* what) {
    std::cout << "Error in " << what << ": " << svo_last_error(ctx) << std::endl;
    std::exit(0);
}

// ---------------------------------------------------------------------------
// File IO at memory speed: positional reads / writes from several threads through rings of pinned chunks.
// ---------------------------------------------------------------------------
bool pread_all(int fd, char* dst, size_t n, off_t off) {
    while (n) {
        const ssize_t r = ::pread(fd, dst, n, off);
        if (r <= 0) return false;
        dst += r; n -= (size_t)r; off += r;
    }
    return true;
}
bool pwrite_all(int fd, const char* src, size_t n, off_t off) {
    while (n) {
        const ssize_t r = ::pwrite(fd, src, n, off);
        if (r <= 0) return false;
        src += r; n -= (size_t)r; off += r;
    }
    return true;
}

// Reads bytes [first, first + total) of `fd` in chunks of `chunk` bytes with `n_threads` readers into a ring of
// pinned buffers and hands the chunks IN ORDER to `consume(ptr, bytes)`. consume's contract is the library's
// double-buffer rule (svo_triangles_append): when it returns, every EARLIER chunk has been copied to the device, so a
// slot is reusable once the chunk after it has been consumed (replaces TriReader's fread loop, TriReader.h:41-79).
// Pinning host memory costs about as much as copying it: the chunks come from a small pool that is allocated ONCE per
// rank and serves the input ring first and the output ring afterwards.
struct PinnedPool {
    std::vector<void*> slot;
    size_t bytes = 0;                         // per slot
    bool init(int n, size_t b) {
        bytes = b;
        for (int i = 0; i < n; i++) { void* p = svo_host_alloc(b); if (!p) return false; slot.push_back(p); }
        return true;
    }
    ~PinnedPool() { for (auto p : slot) svo_host_free(p); }
};

template <class Consume>
bool stream_in(int fd, off_t first, size_t total, size_t chunk_unit, int n_threads, const PinnedPool& pool, Consume consume) {
    if (total == 0) return true;
    const size_t chunk = std::max<size_t>(pool.bytes / chunk_unit, 1) * chunk_unit;      // whole records per chunk
    if (chunk > pool.bytes) return false;
    const size_t n_chunks = (total + chunk - 1) / chunk;
    const int R = (int)std::min<size_t>(n_chunks, pool.slot.size());
    const std::vector<void*>& slot = pool.slot;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<char> ready(n_chunks, 0);
    size_t released = 0;                      // chunks [0, released) no longer occupy their slot
    std::atomic<size_t> next{ 0 };
    std::atomic<bool> failed{ false };
    auto reader = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n_chunks || failed) return;
            {   // slot i % R is free once chunk i - R has been released
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || i < released + (size_t)R; });
            }
            const size_t off = i * chunk, len = std::min(chunk, total - off);
            if (!pread_all(fd, static_cast<char*>(slot[i % R]), len, first + (off_t)off)) failed = true;
            { std::lock_guard<std::mutex> lk(mu); ready[i] = 1; }
            cv.notify_all();
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::min<int>(std::min<int>(n_threads, R), (int)n_chunks); t++) th.emplace_back(reader);
    bool ok = true;
    for (size_t i = 0; i < n_chunks && ok; i++) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return failed || ready[i]; });
        }
        if (failed) { ok = false; break; }
        const size_t off = i * chunk, len = std::min(chunk, total - off);
        if (!consume(slot[i % R], len)) { ok = false; failed = true; }
        { std::lock_guard<std::mutex> lk(mu); released = i; }       // chunk i - 1 has reached the device
        cv.notify_all();
    }
    failed = failed || !ok;
    { std::lock_guard<std::mutex> lk(mu); released = n_chunks; }
    cv.notify_all();
    for (auto& t : th) t.join();
    return ok && !failed;
}

// Streams records [first, first + count) of a context's output (fetch = svo_fetch_nodes / svo_fetch_data) to `fd` at
// byte offset first * rec: device -> pinned chunk (the fetch is synchronous) -> pwrite by a pool of writer threads.
// Positional writes: the order in which chunks reach the file does not matter (replaces writeNode / writeVoxelData,
// octree_io.h:49-66). Chunks stay inside `budget` bytes in total.
bool stream_out(svo_ctx* ctx, int fd, uint64_t first, uint64_t count, uint64_t rec, int (*fetch)(svo_ctx*, uint64_t, uint64_t, void*),
                const PinnedPool& pool, int n_threads, std::string& err) {
    if (count == 0) return true;
    const int R = (int)pool.slot.size();
    n_threads = std::max(1, std::min(n_threads, R - 1));
    const uint64_t per = std::max<uint64_t>(pool.bytes / rec, 1);
    const uint64_t n_chunks = (count + per - 1) / per;
    const std::vector<void*>& slot = pool.slot;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> state(R, 0);             // 0 free, 1 filled (waiting for a writer), 2 being written
    std::vector<uint64_t> slot_first(R, 0), slot_n(R, 0);
    bool done = false, failed = false;
    auto writer = [&]() {
        for (;;) {
            int k = -1;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { for (int i = 0; i < R; i++) if (state[i] == 1) return true; return done || failed; });
                for (int i = 0; i < R; i++) if (state[i] == 1) { k = i; break; }
                if (k < 0) return;
                state[k] = 2;
            }
            const bool ok = pwrite_all(fd, static_cast<const char*>(slot[k]), (size_t)(slot_n[k] * rec), (off_t)(slot_first[k] * rec));
            { std::lock_guard<std::mutex> lk(mu); state[k] = 0; if (!ok) failed = true; }
            cv.notify_all();
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(writer);
    for (uint64_t i = 0; i < n_chunks; i++) {
        int k = -1;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { for (int j = 0; j < R; j++) if (state[j] == 0) return true; return failed; });
            if (failed) break;
            for (int j = 0; j < R; j++) if (state[j] == 0) { k = j; break; }
        }
        const uint64_t f0 = first + i * per, n = std::min<uint64_t>(per, first + count - f0);
        if (fetch(ctx, f0, n, slot[k]) != SVO_OK) { std::lock_guard<std::mutex> lk(mu); failed = true; err = std::string("svo_fetch: ") + svo_last_error(ctx); break; }
        { std::lock_guard<std::mutex> lk(mu); slot_first[k] = f0; slot_n[k] = n; state[k] = 1; }
        cv.notify_all();
    }
    { std::lock_guard<std::mutex> lk(mu); done = true; }
    cv.notify_all();
    for (auto& t : th) t.join();
    if (failed && err.empty()) err = "write error (disk full?)";
    return !failed;
}

int io_threads() {
    if (const char* e = std::getenv("SVO_IO_THREADS")) return std::max(1, std::atoi(e));
    return (int)std::min<unsigned>(std::max(2u, std::thread::hardware_concurrency() / 2), 8u);
}

}  // namespace

int main(int argc, char** argv) {
    WallTimer t_main;
    banner();
    const Options opt = parse(argc, argv);

    // ---- input -----------------------------------------------------------
    WallTimer t_in;
    std::cout << "Parsing tri header " << opt.tri_path << " ..." << std::endl;
    TriHeader hdr;
    if (!read_tri_header(opt.tri_path, hdr)) return 0;
    const std::string tridata = hdr.base + ".tridata";
    if (!file_exists(tridata)) {
        std::cout << "Not all required .tri or .tridata files exist. Please regenerate using tri_convert." << std::endl;
        return 0;
    }
    if (opt.verbose) {
        std::cout << "  base_filename: " << hdr.base << "\n  tri version: " << hdr.version << "\n  geometry only: " << hdr.geometry_only
                  << "\n  n_triangles: " << hdr.n_triangles << "\n  bbox min: " << hdr.bbox_min[0] << " " << hdr.bbox_min[1] << " "
                  << hdr.bbox_min[2] << "\n  bbox max: " << hdr.bbox_max[0] << " " << hdr.bbox_max[1] << " " << hdr.bbox_max[2] << std::endl;
    }
    if (kBinary && !hdr.geometry_only) {
        std::cout << "You're using a .tri file which contains more than just geometry with a geometry-only SVO Builder! "
                     "Regenerate that .tri file using tri_convert_binary." << std::endl;
        return 0;
    }
    if (!kBinary && hdr.geometry_only) {
        std::cout << "You're using a .tri file which contains only geometry with the regular SVO Builder! "
                     "Regenerate that .tri file using tri_convert." << std::endl;
        return 0;
    }

    // Devices. One GPU: only the chosen device is made visible to this process when the user did not restrict the set
    // (initialising all eight devices of a box costs far more than one); with CUDA_VISIBLE_DEVICES given by the user,
    // -g indexes into that set. -gpus N: the first N visible devices.
    const int world = opt.gpus;
    const bool user_set = std::getenv("CUDA_VISIBLE_DEVICES") != nullptr;
    if (!user_set && world == 1) {
        const std::string dev = std::to_string(opt.device);
        setenv("CUDA_VISIBLE_DEVICES", dev.c_str(), 1);
    }
    const int first_device = (world == 1) ? (user_set ? opt.device : 0) : 0;

    // CUDA start-up (driver + one context per device, ~1 s) runs in the background while the file is opened
    std::vector<svo_ctx*> ctx(world, nullptr);
    std::vector<std::future<int>> ctx_ready;
    for (int r = 0; r < world; r++)
        ctx_ready.push_back(std::async(std::launch::async, [&ctx, r, first_device]() { return svo_ctx_create(first_device + r, &ctx[r]); }));

    svo_params prm;
    std::memset(&prm, 0, sizeof prm);
    prm.gridsize = opt.gridsize;
    prm.memory_limit_mb = opt.memory_limit;
    prm.bbox_min0 = hdr.bbox_min[0];
    prm.bbox_max0 = hdr.bbox_max[0];
    prm.payload = kBinary ? 0 : 1;
    prm.generate_levels = opt.levels ? 1 : 0;
    prm.color_mode = opt.color;
    prm.sparseness_limit = opt.sparseness;
    prm.separability = opt.separability;

    const size_t rec_bytes = (size_t)kFloatsPerTri * sizeof(float);
    const size_t tri_bytes = (size_t)hdr.n_triangles * rec_bytes;
    const size_t budget = std::max<size_t>((size_t)opt.memory_limit << 20, 1u << 20);
    const int n_io = io_threads();
    const int in_fd = ::open(tridata.c_str(), O_RDONLY);
    struct stat sb;
    if (in_fd < 0 || ::fstat(in_fd, &sb) != 0 || (size_t)sb.st_size < tri_bytes) {
        std::cout << "Error: " << tridata << " holds fewer than the " << tri_bytes << " bytes the header promises" << std::endl;
        return 0;
    }
    for (int r = 0; r < world; r++) if (ctx_ready[r].get() != SVO_OK) die(nullptr, "svo_ctx_create");

    // Triangle records: the file is read by several threads (pread) into a ring of pinned chunks inside the -l budget
    // and streamed to the device(s) in file order; the whole file is never resident on the host. One pool of pinned
    // chunks per rank (at most 8 x 16 MB, inside the budget), reused for the output.
    const int pool_slots = 8;
    const size_t pool_chunk = std::max<size_t>(std::min<size_t>(budget / (size_t)(pool_slots * world), 16u << 20), 2 * 84);
    std::vector<PinnedPool> pools(world);
    for (int r = 0; r < world; r++) if (!pools[r].init(pool_slots, pool_chunk)) { std::cout << "Error: cannot allocate pinned IO buffers" << std::endl; return 0; }
    const uint64_t per_rank = (hdr.n_triangles + (uint64_t)world - 1) / (uint64_t)world;       // rank r holds slice r of the file
    std::vector<void*> windows(world, nullptr);
    if (world == 1) {
        if (svo_triangles_begin(ctx[0], hdr.n_triangles, kFloatsPerTri) != SVO_OK) die(ctx[0], "svo_triangles_begin");
        const bool ok = stream_in(in_fd, 0, tri_bytes, rec_bytes, n_io, pools[0], [&](void* p, size_t len) {
            return svo_triangles_append(ctx[0], static_cast<const float*>(p), len / rec_bytes) == SVO_OK;
        });
        if (!ok) { std::cout << "Error reading " << tridata << ": " << svo_last_error(ctx[0]) << std::endl; return 0; }
        if (svo_synchronize(ctx[0]) != SVO_OK) die(ctx[0], "svo_synchronize");
    } else {
        for (int r = 0; r < world; r++) {
            if (svo_shard_configure(ctx[r], r, world) != SVO_OK) die(ctx[r], "svo_shard_configure");
            if (svo_shard_slice_create(ctx[r], per_rank, kFloatsPerTri, &windows[r]) != SVO_OK) die(ctx[r], "svo_shard_slice_create");
        }
        for (int r = 0; r < world; r++)
            if (svo_shard_slice_attach(ctx[r], windows.data()) != SVO_OK) die(ctx[r], "svo_shard_slice_attach");
        std::vector<std::future<bool>> up;
        for (int r = 0; r < world; r++) {
            up.push_back(std::async(std::launch::async, [&, r]() {
                const uint64_t lo = std::min<uint64_t>((uint64_t)r * per_rank, hdr.n_triangles), hi = std::min<uint64_t>(lo + per_rank, hdr.n_triangles);
                if (svo_shard_slice_begin(ctx[r], hi - lo) != SVO_OK) return false;
                const bool ok = stream_in(in_fd, (off_t)(lo * rec_bytes), (size_t)(hi - lo) * rec_bytes, rec_bytes, std::max(1, n_io / world), pools[r], [&](void* p, size_t len) {
                    return svo_shard_slice_append(ctx[r], static_cast<const float*>(p), len / rec_bytes) == SVO_OK;
                });
                return ok && svo_synchronize(ctx[r]) == SVO_OK;
            }));
        }
        for (int r = 0; r < world; r++) if (!up[r].get()) { std::cout << "Error reading " << tridata << ": " << svo_last_error(ctx[r]) << std::endl; return 0; }
    }
    ::close(in_fd);
    const double ms_in = t_in.ms();

    // ---- partitioning ----------------------------------------------------
    WallTimer t_part;
    std::cout << "Estimating best partition count ..." << std::endl;
    const uint64_t required = (opt.gridsize * opt.gridsize * opt.gridsize) / 1024 / 1024;
    std::cout << "  to do this in-core I would need " << required << " Mb of system memory" << std::endl;
    const uint64_t P = svo_estimate_partitions(opt.gridsize, opt.memory_limit);
    if (P == 1) std::cout << "  memory limit of " << opt.memory_limit << " Mb allows that" << std::endl;
    else std::cout << "  going to do it in " << P << " partitions of " << required / P << " Mb each." << std::endl;
    std::cout << "Partitioning data into " << P << " partitions ... " << std::flush;

    std::vector<uint64_t> tricounts(P, 0);
    bool have_tricounts = false;
    if (world == 1) {
        uint64_t P_lib = 0;
        if (svo_partition(ctx[0], &prm, &P_lib, tricounts.data(), P) != SVO_OK) die(ctx[0], "svo_partition");
        have_tricounts = true;
    }
    std::cout << "done." << std::endl;
    if (opt.verbose && have_tricounts) {
        for (uint64_t i = 0; i < P; i++) std::cout << "  partition " << i << " - tri_count: " << tricounts[i] << std::endl;
    }
    const double ms_part = t_part.ms();

    // ---- voxelize + build -------------------------------------------------
    // All partitions are voxelized and built in ONE pass on the device(s); the reference's per-partition progress lines
    // (main.cpp:334-352) are printed afterwards, when the per-partition voxel counts exist.
    WallTimer t_vox;
    uint64_t n_voxels = 0, n_nodes = 0, n_data = 0;
    std::vector<uint64_t> voxcounts(P, 0);
    if (world == 1) {
        if (svo_voxelize(ctx[0]) != SVO_OK) die(ctx[0], "svo_voxelize");
        if (svo_build(ctx[0], &n_voxels, &n_nodes, &n_data) != SVO_OK) die(ctx[0], "svo_build");
        if (svo_partition_voxel_counts(ctx[0], voxcounts.data(), P) != SVO_OK) die(ctx[0], "svo_partition_voxel_counts");
    } else {
        // one host thread per device: publish the slice lists, voxelize (staging triangles from the peers' HBM over
        // NVLink), local build, table exchange over peer memory, merged emission. SVO_E_RETRY is answered by every rank
        // in the same step (include/svo_b200.h): all of them repeat the last three calls.
        std::vector<uint64_t> nv(world), nn(world), nd(world);
        std::vector<std::vector<uint64_t>> vc(world, std::vector<uint64_t>(P, 0));
        std::vector<std::future<int>> job;
        for (int r = 0; r < world; r++) {
            job.push_back(std::async(std::launch::async, [&, r]() -> int {
                svo_ctx* c = ctx[r];
                if (svo_shard_slice_publish(c, &prm, hdr.n_triangles) != SVO_OK) return 1;
                if (svo_partition(c, &prm, nullptr, nullptr, 0) != SVO_OK) return 1;
                if (svo_voxelize(c) != SVO_OK) return 1;
                int rc = SVO_E_RETRY;
                for (int attempt = 0; attempt < 4 && rc == SVO_E_RETRY; attempt++) {      // NULL: the library's own table
                    if (svo_shard_count(c, nullptr) != SVO_OK) { rc = 1; break; }
                    if (svo_shard_exchange(c, nullptr) != SVO_OK) { rc = 1; break; }
                    rc = svo_shard_emit(c, nullptr, &nv[r], &nn[r], &nd[r]);
                }
                if (rc == SVO_OK && svo_partition_voxel_counts(c, vc[r].data(), P) != SVO_OK) rc = 1;
                return rc;
            }));
        }
        for (int r = 0; r < world; r++) if (job[r].get() != SVO_OK) die(ctx[r], "the sharded build");
        n_voxels = nv[0]; n_nodes = nn[0]; n_data = nd[0];
        for (int r = 0; r < world; r++) for (uint64_t i = 0; i < P; i++) voxcounts[i] += vc[r][i];
    }
    for (uint64_t i = 0; i < P; i++) {
        if (have_tricounts ? tricounts[i] == 0 : voxcounts[i] == 0) continue;      // main.cpp:330 (without per-partition lists: empty partitions)
        std::cout << "Voxelizing partition " << i << " ..." << std::endl;
        if (opt.verbose && have_tricounts) std::cout << "  reading " << tricounts[i] << " triangles from device list " << i << std::endl;
        if (opt.verbose) std::cout << "  found " << voxcounts[i] << " new voxels." << std::endl;
        std::cout << "Building SVO for partition " << i << " ..." << std::endl;
    }
    const double ms_vox = t_vox.ms();

    // ---- stream the result to disk within the -l budget ---------------------
    WallTimer t_out;
    std::ostringstream name;
    name << hdr.base << opt.gridsize << "_" << P;              // partitioner.cpp:90/141
    const std::string out_base = name.str();
    auto write_file = [&](const std::string& path, uint64_t total, uint64_t rec, bool nodes) {
        const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fd < 0) { std::cout << "Error: cannot open " << path << " for writing" << std::endl; std::exit(0); }
        if (::ftruncate(fd, (off_t)(total * rec)) != 0) { std::cout << "Error: cannot size " << path << std::endl; std::exit(0); }
        std::vector<std::future<std::string>> w;
        for (int r = 0; r < world; r++) {
            w.push_back(std::async(std::launch::async, [&, r]() -> std::string {
                uint64_t lo = 0, hi = total;
                if (world > 1) {
                    uint64_t nlo, nhi, dlo, dhi;
                    if (svo_shard_ranges(ctx[r], &nlo, &nhi, &dlo, &dhi) != SVO_OK) return svo_last_error(ctx[r]);
                    lo = nodes ? nlo : dlo; hi = nodes ? nhi : dhi;
                }
                std::string err;
                stream_out(ctx[r], fd, lo, hi - lo, rec, nodes ? svo_fetch_nodes : svo_fetch_data, pools[r], std::max(1, n_io / world), err);
                return err;
            }));
        }
        for (int r = 0; r < world; r++) { const std::string e = w[r].get(); if (!e.empty()) { std::cout << "Error writing " << path << ": " << e << std::endl; std::exit(0); } }
        if (::close(fd) != 0) { std::cout << "Error closing " << path << std::endl; std::exit(0); }
    };
    write_file(out_base + ".octreenodes", n_nodes, SVO_NODE_BYTES, true);
    write_file(out_base + ".octreedata", n_data, SVO_DATA_BYTES, false);
    {
        std::ofstream h((out_base + ".octree").c_str());        // octree_io.h:74-83
        h << "#octreeheader 1\n" << "gridlength " << opt.gridsize << "\n" << "n_nodes " << n_nodes << "\n" << "n_data " << n_data << "\nEND\n";
        h.flush();
        if (!h) { std::cout << "Error writing " << out_base << ".octree" << std::endl; return 0; }
    }
    const double ms_out = t_out.ms();
    std::cout << "done" << std::endl;
    std::cout << "Total amount of voxels: " << n_voxels << std::endl;

    svo_stats st;
    svo_get_stats(ctx[0], &st);
    // same sections as the reference's printTimerInfo (main.cpp:219-241); algorithm times are CUDA-event device times
    // (rank 0's with -gpus > 1)
    std::cout << "Total MAIN time      : " << t_main.ms() << " ms." << std::endl;
    std::cout << "PARTITIONING\n  Total time\t\t: " << ms_part + ms_in << " ms.\n  IO IN time\t\t: " << ms_in
              << " ms.\n  algorithm time\t: " << st.ms_partition << " ms. (device)\n  upload time\t\t: " << st.ms_upload << " ms. (device)" << std::endl;
    std::cout << "VOXELIZING\n  Total time\t\t: " << ms_vox << " ms. (voxelizing + SVO building, host wall clock)\n  algorithm time\t: "
              << st.ms_voxelize << " ms. (device)" << std::endl;
    std::cout << "SVO BUILDING\n  algorithm time\t: " << st.ms_build << " ms. (device)\n  IO OUT time\t\t: " << ms_out << " ms." << std::endl;
    if (opt.verbose) {
        std::cout << "  pairs: " << st.n_pairs << " (small " << st.n_small << ", medium " << st.n_medium << ", large " << st.n_large << ")\n"
                  << "  nodes: " << st.n_nodes << "  data: " << st.n_data << "  kernel launches: " << st.kernel_launches
                  << "  devices: " << world << "  io threads: " << n_io << std::endl;
    }
    for (int r = 0; r < world; r++) svo_ctx_destroy(ctx[r]);
    return 0;
}
