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

#include <algorithm>
#include <chrono>
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

    // CUDA start-up (driver + context, ~1 s) runs in the background while the triangle file is read. Only the
    // chosen GPU is made visible to this process: initialising all eight devices of a box costs far more.
    if (!std::getenv("CUDA_VISIBLE_DEVICES")) {
        const std::string dev = std::to_string(opt.device);
        setenv("CUDA_VISIBLE_DEVICES", dev.c_str(), 1);
    }
    const int device_index = std::getenv("SVO_KEEP_DEVICE_INDEX") ? opt.device : 0;
    svo_ctx* ctx = nullptr;
    std::future<int> ctx_ready = std::async(std::launch::async, [&ctx, device_index]() { return svo_ctx_create(device_index, &ctx); });

    // Triangle records: the file is read in chunks into two alternating pinned buffers and streamed to the device, the
    // copy of one chunk overlapping the fread of the next (replaces TriReader's 8192-triangle fread loop,
    // TriReader.h:41-79). The two chunks stay inside the -l budget; the whole file is never resident on the host.
    const size_t rec_bytes = (size_t)kFloatsPerTri * sizeof(float);
    const size_t tri_bytes = (size_t)hdr.n_triangles * rec_bytes;
    {
        const size_t in_budget = std::max<size_t>((size_t)opt.memory_limit << 20, 1u << 20);
        const size_t chunk_tris = std::max<size_t>(std::min<size_t>(in_budget / 4, 64u << 20) / rec_bytes, 1);
        FILE* f = std::fopen(tridata.c_str(), "rb");
        // while the CUDA context comes up (~1 s), read the head of the file into ordinary memory (within the budget)
        const size_t head_cap = std::min<size_t>(tri_bytes, (in_budget / 2) / rec_bytes * rec_bytes);
        char* head = static_cast<char*>(std::malloc(head_cap ? head_cap : 1));
        size_t head_have = 0;
        while (f && head && head_have < head_cap && ctx_ready.wait_for(std::chrono::seconds(0)) != std::future_status::ready) {
            const size_t want = std::min<size_t>(head_cap - head_have, (8u << 20) / rec_bytes * rec_bytes + rec_bytes);
            const size_t r = std::fread(head + head_have, 1, want, f);
            if (r != want) { head_have += r; break; }
            head_have += r;
        }
        head_have = head_have / rec_bytes * rec_bytes;
        if (f) std::fseek(f, (long)head_have, SEEK_SET);
        if (ctx_ready.get() != SVO_OK) die(nullptr, "svo_ctx_create");
        void* in_chunk[2] = { svo_host_alloc(chunk_tris * rec_bytes), svo_host_alloc(chunk_tris * rec_bytes) };
        if (!in_chunk[0] || !in_chunk[1]) { std::cout << "Error: cannot allocate pinned input buffers" << std::endl; return 0; }
        if (svo_triangles_begin(ctx, hdr.n_triangles, kFloatsPerTri) != SVO_OK) die(ctx, "svo_triangles_begin");
        size_t got = 0;
        if (head_have) {
            if (svo_triangles_append(ctx, reinterpret_cast<const float*>(head), head_have / rec_bytes) != SVO_OK) die(ctx, "svo_triangles_append");
            if (svo_synchronize(ctx) != SVO_OK) die(ctx, "svo_synchronize");      // `head` is pageable and freed right away
            got = head_have;
        }
        std::free(head);
        int slot = 0;
        while (f && got < tri_bytes) {
            const size_t want = std::min<size_t>(tri_bytes - got, chunk_tris * rec_bytes);
            size_t have = 0;
            while (have < want) {
                const size_t r = std::fread(static_cast<char*>(in_chunk[slot]) + have, 1, want - have, f);
                if (r == 0) break;
                have += r;
            }
            if (have != want) break;
            if (svo_triangles_append(ctx, static_cast<const float*>(in_chunk[slot]), want / rec_bytes) != SVO_OK) die(ctx, "svo_triangles_append");
            got += want;
            slot ^= 1;
        }
        if (f) std::fclose(f);
        if (got != tri_bytes) { std::cout << "Error: " << tridata << " holds fewer than the " << tri_bytes << " bytes the header promises" << std::endl; return 0; }
        if (svo_synchronize(ctx) != SVO_OK) die(ctx, "svo_synchronize");
        svo_host_free(in_chunk[0]);
        svo_host_free(in_chunk[1]);
    }
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

    std::vector<uint64_t> tricounts(P, 0);
    uint64_t P_lib = 0;
    if (svo_partition(ctx, &prm, &P_lib, tricounts.data(), P) != SVO_OK) die(ctx, "svo_partition");
    std::cout << "done." << std::endl;
    if (opt.verbose) {
        for (uint64_t i = 0; i < P; i++) std::cout << "  partition " << i << " - tri_count: " << tricounts[i] << std::endl;
    }
    const double ms_part = t_part.ms();

    // ---- voxelize + build -------------------------------------------------
    WallTimer t_vox;
    for (uint64_t i = 0; i < P; i++) {
        if (tricounts[i] == 0) continue;                       // main.cpp:330
        std::cout << "Voxelizing partition " << i << " ..." << std::endl;
        if (opt.verbose) std::cout << "  reading " << tricounts[i] << " triangles from device list " << i << std::endl;
        std::cout << "Building SVO for partition " << i << " ..." << std::endl;
    }
    if (svo_voxelize(ctx) != SVO_OK) die(ctx, "svo_voxelize");
    uint64_t n_voxels = 0, n_nodes = 0, n_data = 0;
    if (svo_build(ctx, &n_voxels, &n_nodes, &n_data) != SVO_OK) die(ctx, "svo_build");
    const double ms_vox = t_vox.ms();

    // ---- stream the result to disk within the -l budget ---------------------
    WallTimer t_out;
    std::ostringstream name;
    name << hdr.base << opt.gridsize << "_" << P;              // partitioner.cpp:90/141
    const std::string out_base = name.str();
    const size_t budget = std::max<size_t>((size_t)opt.memory_limit << 20, 1u << 20);
    const size_t chunk_bytes = std::min<size_t>(budget / 2, 128u << 20);          // two chunks in flight stay inside the -l budget
    void* chunk[2] = { svo_host_alloc(chunk_bytes), svo_host_alloc(chunk_bytes) };
    if (!chunk[0] || !chunk[1]) { std::cout << "Error: cannot allocate pinned output buffers" << std::endl; return 0; }
    // device -> pinned chunk (svo_fetch_*) overlaps with fwrite of the previous chunk
    auto stream_out = [&](const std::string& path, uint64_t count, uint64_t rec, int (*fetch)(svo_ctx*, uint64_t, uint64_t, void*)) {
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f) { std::cout << "Error: cannot open " << path << " for writing" << std::endl; std::exit(0); }
        const uint64_t per = chunk_bytes / rec;
        std::future<void> writing[2];
        int slot = 0;
        for (uint64_t first = 0; first < count; first += per, slot ^= 1) {
            const uint64_t n = std::min<uint64_t>(per, count - first);
            if (writing[slot].valid()) writing[slot].get();
            if (fetch(ctx, first, n, chunk[slot]) != SVO_OK) die(ctx, "svo_fetch");
            void* src = chunk[slot];
            writing[slot] = std::async(std::launch::async, [f, src, rec, n]() { std::fwrite(src, rec, n, f); });
            if (writing[slot ^ 1].valid()) writing[slot ^ 1].get();          // keep the file writes in order
        }
        for (auto& w : writing) if (w.valid()) w.get();
        std::fclose(f);
    };
    stream_out(out_base + ".octreenodes", n_nodes, SVO_NODE_BYTES, svo_fetch_nodes);
    stream_out(out_base + ".octreedata", n_data, SVO_DATA_BYTES, svo_fetch_data);
    {
        std::ofstream h((out_base + ".octree").c_str());        // octree_io.h:74-83
        h << "#octreeheader 1\n" << "gridlength " << opt.gridsize << "\n" << "n_nodes " << n_nodes << "\n" << "n_data " << n_data << "\nEND\n";
    }
    const double ms_out = t_out.ms();
    std::cout << "done" << std::endl;
    std::cout << "Total amount of voxels: " << n_voxels << std::endl;

    svo_stats st;
    svo_get_stats(ctx, &st);
    // same sections as the reference's printTimerInfo (main.cpp:219-241); algorithm times are CUDA-event device times
    std::cout << "Total MAIN time      : " << t_main.ms() << " ms." << std::endl;
    std::cout << "PARTITIONING\n  Total time\t\t: " << ms_part + ms_in << " ms.\n  IO IN time\t\t: " << ms_in
              << " ms.\n  algorithm time\t: " << st.ms_partition << " ms. (device)\n  upload time\t\t: " << st.ms_upload << " ms. (device)" << std::endl;
    std::cout << "VOXELIZING\n  Total time\t\t: " << ms_vox << " ms. (voxelizing + SVO building, host wall clock)\n  algorithm time\t: "
              << st.ms_voxelize << " ms. (device)" << std::endl;
    std::cout << "SVO BUILDING\n  algorithm time\t: " << st.ms_build << " ms. (device)\n  IO OUT time\t\t: " << ms_out << " ms." << std::endl;
    if (opt.verbose) {
        std::cout << "  pairs: " << st.n_pairs << " (small " << st.n_small << ", medium " << st.n_medium << ", large " << st.n_large << ")\n"
                  << "  nodes: " << st.n_nodes << "  data: " << st.n_data << "  kernel launches: " << st.kernel_launches << std::endl;
    }
    svo_host_free(chunk[0]);
    svo_host_free(chunk[1]);
    svo_ctx_destroy(ctx);
    return 0;
}
