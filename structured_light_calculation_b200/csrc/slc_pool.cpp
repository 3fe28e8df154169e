// slc_pool.cpp -- several GPUs behind one call (include/slcalc_b200.h "slc_pool").
//
// The frame sets of a DynaFrame sequence are independent in the north-star definition, so the
// multi-GPU form of the path is a partition, not a collective: one context + one feeder thread
// per GPU, calibration replicated, frame sets handed out ON DEMAND (max_batch at a time from one
// atomic counter) so that every GPU keeps its upload / kernel / download slots full and a GPU
// behind a faster host link takes more of the batch.  This file is host C++ on nothing but the
// public C ABI -- a feeder does exactly what a single-GPU caller would do with its own context
// (slc_submit_host_ex / slc_wait over the context's stream slots).  What it replaces in the
// reference: the serial frame loop of main.cpp:42-45 / CCalculation.cpp:221, which a reference
// user would have to thread by hand.
#include "../../include/slcalc_b200.h"

#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

struct slc_pool {
    std::vector<slc_context*> ctx;
    std::vector<std::thread> feeders;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::function<int(int)> job;          // member index -> status
    unsigned long long generation = 0;    // bumped per job
    int remaining = 0;
    bool quit = false;
    std::vector<int> status;
    std::vector<int> shares;              // frame sets each member took in the last host call
    int chunk = 1, slots = 1;             // max_batch / num_slots of every member
    std::string err;
};

namespace {

thread_local std::string g_pool_create_error;

int pool_fail(slc_pool* pool, int status, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (pool) pool->err = buf; else g_pool_create_error = buf;
    return status;
}

void feeder_main(slc_pool* pool, int member)
{
    unsigned long long seen = 0;
    for (;;) {
        std::function<int(int)> job;
        {
            std::unique_lock<std::mutex> lk(pool->mu);
            pool->cv_job.wait(lk, [&] { return pool->quit || pool->generation != seen; });
            if (pool->quit) return;
            seen = pool->generation;
            job = pool->job;
        }
        const int st = job(member);
        {
            std::lock_guard<std::mutex> lk(pool->mu);
            pool->status[(size_t)member] = st;
            pool->remaining--;
        }
        pool->cv_done.notify_all();
    }
}

// every member runs job(member) on its own feeder thread; first failure wins
int run_all(slc_pool* pool, std::function<int(int)> job)
{
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        pool->job = std::move(job);
        pool->remaining = (int)pool->ctx.size();
        pool->generation++;
    }
    pool->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(pool->mu);
        pool->cv_done.wait(lk, [&] { return pool->remaining == 0; });
    }
    for (size_t i = 0; i < pool->ctx.size(); i++)
        if (pool->status[i] != SLC_OK)
            return pool_fail(pool, pool->status[i], "member %zu: %s", i, slc_last_error(pool->ctx[i]));
    return SLC_OK;
}

}  // namespace

extern "C" {

int slc_shard_range(int64_t n_items, int32_t index, int32_t n_shards, int64_t* lo, int64_t* hi)
{
    if (n_items < 0 || n_shards < 1 || index < 0 || index >= n_shards || !lo || !hi) return SLC_ERR_INVALID_ARG;
    const int64_t base = n_items / n_shards, rem = n_items % n_shards;
    *lo = index * base + (index < rem ? index : rem);
    *hi = *lo + base + (index < rem ? 1 : 0);
    return SLC_OK;
}

int slc_pool_create(const slc_config* cfg, const int32_t* devices, int32_t n_devices, slc_pool** out)
{
    if (!cfg || !devices || !out || n_devices < 1 || n_devices > 64)
        return pool_fail(nullptr, SLC_ERR_INVALID_ARG, "slc_pool_create: NULL argument or n_devices outside 1..64");
    *out = nullptr;
    slc_pool* pool = new (std::nothrow) slc_pool();
    if (!pool) return pool_fail(nullptr, SLC_ERR_OUT_OF_MEMORY, "host allocation failed");
    try {
        pool->status.assign((size_t)n_devices, SLC_OK);
        pool->shares.assign((size_t)n_devices, 0);
        pool->chunk = cfg->max_batch;
        pool->slots = cfg->num_slots;
        for (int i = 0; i < n_devices; i++) {
            slc_config c = *cfg;
            c.device = devices[i];
            slc_context* ctx = nullptr;
            const int st = slc_create(&c, &ctx);
            if (st != SLC_OK) {
                pool_fail(nullptr, st, "member %d (device %d): %s", i, devices[i], slc_last_error(nullptr));
                slc_pool_destroy(pool);
                return st;
            }
            pool->ctx.push_back(ctx);
        }
        for (int i = 0; i < n_devices; i++) pool->feeders.emplace_back(feeder_main, pool, i);
    } catch (const std::exception& ex) {      // no exception leaves the C ABI
        pool_fail(nullptr, SLC_ERR_OUT_OF_MEMORY, "slc_pool_create: %s", ex.what());
        slc_pool_destroy(pool);
        return SLC_ERR_OUT_OF_MEMORY;
    }
    *out = pool;
    return SLC_OK;
}

void slc_pool_destroy(slc_pool* pool)
{
    if (!pool) return;
    {
        std::lock_guard<std::mutex> lk(pool->mu);
        pool->quit = true;
    }
    pool->cv_job.notify_all();
    for (auto& t : pool->feeders)
        if (t.joinable()) t.join();
    for (slc_context* c : pool->ctx) slc_destroy(c);
    delete pool;
}

const char* slc_pool_last_error(const slc_pool* pool) { return pool ? pool->err.c_str() : g_pool_create_error.c_str(); }

int slc_pool_size(const slc_pool* pool) { return pool ? (int)pool->ctx.size() : 0; }

slc_context* slc_pool_context(slc_pool* pool, int32_t member)
{
    if (!pool || member < 0 || member >= (int)pool->ctx.size()) return nullptr;
    return pool->ctx[(size_t)member];
}

int slc_pool_set_calibration(slc_pool* pool, const double cam[9], const double pro[9], const double R[9], const double T[3])
{
    if (!pool) return SLC_ERR_INVALID_ARG;
    for (size_t i = 0; i < pool->ctx.size(); i++) {
        const int st = slc_set_calibration(pool->ctx[i], cam, pro, R, T);
        if (st != SLC_OK) return pool_fail(pool, st, "member %zu: %s", i, slc_last_error(pool->ctx[i]));
    }
    return SLC_OK;
}

int slc_pool_set_gray_lut(slc_pool* pool, const int16_t* gray2bin, int32_t n)
{
    if (!pool) return SLC_ERR_INVALID_ARG;
    for (size_t i = 0; i < pool->ctx.size(); i++) {
        const int st = slc_set_gray_lut(pool->ctx[i], gray2bin, n);
        if (st != SLC_OK) return pool_fail(pool, st, "member %zu: %s", i, slc_last_error(pool->ctx[i]));
    }
    return SLC_OK;
}

int slc_pool_reconstruct_host(slc_pool* pool, const uint8_t* h_stack, int32_t n_stacks, const slc_result* h_out)
{
    if (!pool) return SLC_ERR_INVALID_ARG;
    if (n_stacks < 0) return pool_fail(pool, SLC_ERR_INVALID_ARG, "n_stacks < 0");
    for (int& v : pool->shares) v = 0;
    if (n_stacks == 0) return SLC_OK;
    if (!h_stack || !h_out) return pool_fail(pool, SLC_ERR_INVALID_ARG, "NULL buffer");
    slc_info info;
    int st = slc_get_info(pool->ctx[0], &info);
    if (st != SLC_OK) return pool_fail(pool, st, "member 0: %s", slc_last_error(pool->ctx[0]));
    const size_t npx = (size_t)info.pixels, sb = (size_t)info.stack_bytes, bb = (npx + 7) / 8;
    const int chunk = pool->chunk, slots = pool->slots;
    const slc_result r = *h_out;
    std::atomic<int> next{0};             // first frame set nobody has taken yet
    std::atomic<bool> stop{false};        // a member failed: the others stop taking work
    return run_all(pool, [&, npx, sb, bb, chunk, slots, r](int i) -> int {
        slc_context* c = pool->ctx[(size_t)i];
        std::vector<char> busy((size_t)slots, 0);
        int rc = SLC_OK, slot = 0, taken = 0;
        while (rc == SLC_OK && !stop.load(std::memory_order_relaxed)) {
            if (busy[(size_t)slot]) {                       // the chunk this slot ran `slots` chunks ago
                rc = slc_wait(c, slot);
                busy[(size_t)slot] = 0;
                if (rc != SLC_OK) break;
            }
            const int lo = next.fetch_add(chunk, std::memory_order_relaxed);
            if (lo >= n_stacks) break;
            const int n = (n_stacks - lo) < chunk ? (n_stacks - lo) : chunk;
            slc_result o = r;                               // the planes of frame sets lo .. lo + n - 1
            const size_t d = (size_t)lo;
            if (r.xyzw) o.xyzw = r.xyzw + d * npx * 4;
            if (r.mask) o.mask = r.mask + d * npx;
            if (r.depth) o.depth = r.depth + d * npx;
            if (r.mask_bits) o.mask_bits = r.mask_bits + d * bb;
            if (r.points) o.points = r.points + d * (size_t)r.point_stride * 3;
            if (r.n_points) o.n_points = r.n_points + d;
            rc = slc_submit_host_ex(c, slot, h_stack + d * sb, n, &o);
            if (rc != SLC_OK) break;
            busy[(size_t)slot] = 1;
            taken += n;
            slot = (slot + 1) % slots;
        }
        if (rc != SLC_OK) stop.store(true, std::memory_order_relaxed);
        for (int k = 0; k < slots; k++) {                   // drain, oldest first
            const int q = (slot + k) % slots;
            if (!busy[(size_t)q]) continue;
            const int rc2 = slc_wait(c, q);
            if (rc == SLC_OK) rc = rc2;
        }
        pool->shares[(size_t)i] = taken;
        return rc;
    });
}

int slc_pool_last_shares(const slc_pool* pool, int32_t* frame_sets_per_member, int32_t n_members)
{
    if (!pool || !frame_sets_per_member || n_members != (int)pool->ctx.size()) return SLC_ERR_INVALID_ARG;
    for (int i = 0; i < n_members; i++) frame_sets_per_member[i] = pool->shares[(size_t)i];
    return SLC_OK;
}

int slc_pool_reconstruct_device(slc_pool* pool, const uint8_t* const* d_stack, const int32_t* n_stacks,
                                const slc_result* d_out)
{
    if (!pool) return SLC_ERR_INVALID_ARG;
    if (!d_stack || !n_stacks || !d_out) return pool_fail(pool, SLC_ERR_INVALID_ARG, "NULL argument");
    return run_all(pool, [=](int i) -> int {
        if (n_stacks[i] == 0) return SLC_OK;
        slc_context* c = pool->ctx[(size_t)i];
        const int st = slc_reconstruct_device_ex(c, d_stack[i], n_stacks[i], &d_out[i], nullptr);
        return st != SLC_OK ? st : slc_synchronize(c);
    });
}

}  // extern "C"
