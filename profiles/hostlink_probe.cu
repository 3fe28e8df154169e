// hostlink_probe.cu -- what the host<->device path of a box can carry, isolated from the library.
//
// The streamed first-frame path (slc_reconstruct_host / slc_pool_reconstruct_host) moves 50.7 MB up
// and 39.2 MB down per 1920x1200 frame set against 13.7 us of kernel time, so its ceiling is the
// host link.  This probe measures that ceiling with nothing but cudaMemcpyAsync on pinned buffers:
// H2D only, D2H only and both at once; every GPU alone, pairs, quads and all together; three kinds of
// host memory (cudaHostAlloc, write-combined, 2 MB-page mmap + cudaHostRegister); one process with a
// thread per GPU, or (driven by hostlink_probe.py) one process per GPU started at a common wall-clock
// time.  A CPU memcpy sweep gives the host DRAM bandwidth beside it.  One JSON object per line.
//
// Build: make -C structured_light_calculation_b200 bin/hostlink_probe
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

using Clock = std::chrono::steady_clock;

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            std::exit(2);                                                                     \
        }                                                                                     \
    } while (0)

enum Mem { kPinned = 0, kWC = 1, kHuge = 2, kMemKinds = 3 };
const char* kMemName[kMemKinds] = {"pinned", "write_combined", "hugepage_registered"};
enum Dir { kH2D = 0, kD2H = 1, kBidir = 2 };
const char* kDirName[3] = {"h2d", "d2h", "bidir"};

struct Dev {
    int id = 0;
    cudaStream_t up = nullptr, dn = nullptr;
    void *d_in = nullptr, *d_out = nullptr;
    void* h_up[kMemKinds] = {};
    void* h_dn[kMemKinds] = {};
};

size_t g_bytes = 64u << 20;
int g_reps = 24;
bool g_huge_ok = false;

void* alloc_host(Mem kind, size_t bytes)
{
    void* p = nullptr;
    if (kind == kPinned) {
        CK(cudaHostAlloc(&p, bytes, cudaHostAllocPortable));
    } else if (kind == kWC) {
        CK(cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocWriteCombined));
    } else {
        // 2 MB pages: explicit hugetlb pages if the box has a pool, else transparent huge pages
        p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB, -1, 0);
        if (p == MAP_FAILED) {
            p = mmap(nullptr, bytes + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (p == MAP_FAILED) return nullptr;
            p = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(p) + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
            madvise(p, bytes, MADV_HUGEPAGE);
        }
        std::memset(p, 1, bytes);
        if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        g_huge_ok = true;
        return p;
    }
    std::memset(p, 1, bytes);
    return p;
}

long anon_huge_kb()
{
    FILE* f = std::fopen("/proc/self/smaps_rollup", "r");
    if (!f) return -1;
    char line[256];
    long kb = -1;
    while (std::fgets(line, sizeof line, f))
        if (std::sscanf(line, "AnonHugePages: %ld kB", &kb) == 1) break;
    std::fclose(f);
    return kb;
}

struct Result { double t0 = 0, t1 = 0; };

// every thread of the experiment spins on the same start flag, so the copies begin together
void run_one(Dev& dv, Dir dir, Mem mem, std::atomic<int>& ready, std::atomic<int>& go, Result& r, double start_at)
{
    CK(cudaSetDevice(dv.id));
    auto issue = [&](int n) {
        for (int i = 0; i < n; i++) {
            if (dir != kD2H) CK(cudaMemcpyAsync(dv.d_in, dv.h_up[mem], g_bytes, cudaMemcpyHostToDevice, dv.up));
            if (dir != kH2D) CK(cudaMemcpyAsync(dv.h_dn[mem], dv.d_out, g_bytes, cudaMemcpyDeviceToHost, dv.dn));
        }
    };
    issue(2);
    CK(cudaStreamSynchronize(dv.up));
    CK(cudaStreamSynchronize(dv.dn));
    ready.fetch_add(1);
    while (go.load(std::memory_order_acquire) == 0) {}
    if (start_at > 0) {
        while (std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count() < start_at) {}
    }
    const auto t0 = Clock::now();
    issue(g_reps);
    CK(cudaStreamSynchronize(dv.up));
    CK(cudaStreamSynchronize(dv.dn));
    const auto t1 = Clock::now();
    r.t0 = std::chrono::duration<double>(t0.time_since_epoch()).count();
    r.t1 = std::chrono::duration<double>(t1.time_since_epoch()).count();
}

void experiment(std::vector<Dev>& devs, const std::vector<int>& set, Dir dir, Mem mem, const char* mode, double start_at)
{
    for (int d : set)
        if (!devs[(size_t)d].h_up[mem] || !devs[(size_t)d].h_dn[mem]) return;   // this memory kind is not available
    std::atomic<int> ready{0}, go{0};
    std::vector<Result> res(set.size());
    std::vector<std::thread> th;
    for (size_t k = 0; k < set.size(); k++)
        th.emplace_back(run_one, std::ref(devs[(size_t)set[k]]), dir, mem, std::ref(ready), std::ref(go), std::ref(res[k]), start_at);
    while (ready.load() < (int)set.size()) std::this_thread::yield();
    go.store(1, std::memory_order_release);
    for (auto& t : th) t.join();
    const double per_dir = (double)g_bytes * g_reps;
    const double per_gpu_bytes = per_dir * (dir == kBidir ? 2 : 1);
    double tmin = 1e300, tmax = 0;
    std::string per = "[";
    for (size_t k = 0; k < set.size(); k++) {
        tmin = std::min(tmin, res[k].t0);
        tmax = std::max(tmax, res[k].t1);
        char buf[64];
        std::snprintf(buf, sizeof buf, "%s%.2f", k ? ", " : "", per_gpu_bytes / (res[k].t1 - res[k].t0) / 1e9);
        per += buf;
    }
    per += "]";
    std::string s = "[";
    for (size_t k = 0; k < set.size(); k++) s += (k ? ", " : "") + std::to_string(devs[(size_t)set[k]].id);
    s += "]";
    std::printf("{\"kind\": \"dma\", \"mode\": \"%s\", \"gpus\": %s, \"n\": %zu, \"dir\": \"%s\", \"mem\": \"%s\", "
                "\"mb_per_copy\": %zu, \"copies_per_dir\": %d, \"per_gpu_gbs\": %s, \"aggregate_gbs\": %.2f, "
                "\"seconds\": %.4f}\n",
                mode, s.c_str(), set.size(), kDirName[dir], kMemName[mem], g_bytes >> 20, g_reps, per.c_str(),
                per_gpu_bytes * set.size() / (tmax - tmin) / 1e9, tmax - tmin);
    std::fflush(stdout);
}

void cpu_stream(int threads)
{
    const size_t bytes = 256u << 20;
    std::vector<char*> a((size_t)threads), b((size_t)threads);
    for (int t = 0; t < threads; t++) {
        a[(size_t)t] = static_cast<char*>(std::malloc(bytes));
        b[(size_t)t] = static_cast<char*>(std::malloc(bytes));
        std::memset(a[(size_t)t], 1, bytes);
        std::memset(b[(size_t)t], 2, bytes);
    }
    std::atomic<int> ready{0}, go{0};
    std::vector<double> t1((size_t)threads);
    std::vector<std::thread> th;
    const int reps = 4;
    Clock::time_point t0;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t] {
            ready.fetch_add(1);
            while (go.load(std::memory_order_acquire) == 0) {}
            for (int r = 0; r < reps; r++) std::memcpy(b[(size_t)t], a[(size_t)t], bytes);
            t1[(size_t)t] = std::chrono::duration<double>(Clock::now().time_since_epoch()).count();
        });
    while (ready.load() < threads) std::this_thread::yield();
    t0 = Clock::now();
    go.store(1, std::memory_order_release);
    for (auto& t : th) t.join();
    const double start = std::chrono::duration<double>(t0.time_since_epoch()).count();
    const double end = *std::max_element(t1.begin(), t1.end());
    std::printf("{\"kind\": \"cpu_memcpy\", \"threads\": %d, \"mb_per_copy\": 256, \"copies\": %d, "
                "\"read_plus_write_gbs\": %.2f}\n", threads, reps, 2.0 * bytes * reps * threads / (end - start) / 1e9);
    std::fflush(stdout);
    for (int t = 0; t < threads; t++) { std::free(a[(size_t)t]); std::free(b[(size_t)t]); }
}

std::vector<int> parse_list(const char* s)
{
    std::vector<int> v;
    for (const char* p = s; *p;) {
        v.push_back(std::atoi(p));
        while (*p && *p != ',') p++;
        if (*p == ',') p++;
    }
    return v;
}

}  // namespace

int main(int argc, char** argv)
{
    std::string plan = "full";
    std::vector<int> only;
    double start_at = 0, spacing = 0.75;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--mb" && i + 1 < argc) g_bytes = (size_t)std::atoi(argv[++i]) << 20;
        else if (a == "--reps" && i + 1 < argc) g_reps = std::atoi(argv[++i]);
        else if (a == "--plan" && i + 1 < argc) plan = argv[++i];          // full | process (one GPU, timed starts)
        else if (a == "--devices" && i + 1 < argc) only = parse_list(argv[++i]);
        else if (a == "--start-at" && i + 1 < argc) start_at = std::atof(argv[++i]);
        else if (a == "--spacing" && i + 1 < argc) spacing = std::atof(argv[++i]);
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 1; }
    }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { std::fprintf(stderr, "no CUDA device\n"); return 3; }
    std::vector<int> ids = only;
    if (ids.empty()) for (int d = 0; d < n_dev; d++) ids.push_back(d);
    const bool all_kinds = plan == "full";
    std::vector<Dev> devs(ids.size());
    for (size_t k = 0; k < ids.size(); k++) {
        Dev& dv = devs[k];
        dv.id = ids[k];
        CK(cudaSetDevice(dv.id));
        CK(cudaStreamCreateWithFlags(&dv.up, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&dv.dn, cudaStreamNonBlocking));
        CK(cudaMalloc(&dv.d_in, g_bytes));
        CK(cudaMalloc(&dv.d_out, g_bytes));
        CK(cudaMemset(dv.d_out, 3, g_bytes));
        for (int m = 0; m < (all_kinds ? (int)kMemKinds : 1); m++) {
            dv.h_up[m] = alloc_host((Mem)m, g_bytes);
            dv.h_dn[m] = alloc_host((Mem)m, g_bytes);
        }
    }
    {
        cudaDeviceProp pr;
        CK(cudaGetDeviceProperties(&pr, ids[0]));
        std::printf("{\"kind\": \"header\", \"plan\": \"%s\", \"devices\": %zu, \"gpu\": \"%s\", \"host_threads\": %u, "
                    "\"hugepage_buffers\": %s, \"anon_huge_kb\": %ld}\n", plan.c_str(), ids.size(), pr.name,
                    std::thread::hardware_concurrency(), g_huge_ok ? "true" : "false", anon_huge_kb());
    }
    const int n = (int)devs.size();
    if (plan == "process") {
        // one process per GPU (hostlink_probe.py starts them): experiment k begins at start_at + k*spacing
        int k = 0;
        for (int dir = 0; dir < 3; dir++, k++)
            experiment(devs, {0}, (Dir)dir, kPinned, "processes", start_at > 0 ? start_at + k * spacing : 0);
        return 0;
    }
    auto range = [](int lo, int hi, int step = 1) { std::vector<int> v; for (int i = lo; i < hi; i += step) v.push_back(i); return v; };
    std::vector<std::vector<int>> sets;
    for (int d = 0; d < n; d++) sets.push_back({d});
    if (n >= 2) sets.push_back({0, 1});
    if (n >= 4) { sets.push_back({0, n / 2}); sets.push_back({0, n - 1}); sets.push_back(range(0, 4)); }
    if (n >= 8) { sets.push_back(range(4, 8)); sets.push_back(range(0, 8, 2)); sets.push_back(range(0, 8)); }
    for (const auto& s : sets)
        for (int dir = 0; dir < 3; dir++) experiment(devs, s, (Dir)dir, kPinned, "threads", 0);
    for (int m = 1; m < kMemKinds; m++)
        for (const auto& s : {std::vector<int>{0}, range(0, n)})
            for (int dir = 0; dir < 3; dir++) experiment(devs, s, (Dir)dir, (Mem)m, "threads", 0);
    // transfer size: does the aggregate depend on the size of a copy? (same bytes in flight, 4 MB copies)
    {
        const size_t keep = g_bytes; const int keep_reps = g_reps;
        g_bytes = 4u << 20; g_reps = keep_reps * 16;
        for (const auto& s : {std::vector<int>{0}, range(0, n)}) experiment(devs, s, kBidir, kPinned, "threads", 0);
        g_bytes = keep; g_reps = keep_reps;
    }
    const int hc = (int)std::thread::hardware_concurrency();
    for (int t : {1, 4, 8, 16, 32, 64})
        if (t <= hc) cpu_stream(t);
    return 0;
}
