// fb_cache.cuh — caching device allocator, one free list per CONTEXT.  cudaMalloc/cudaFree cost ~0.1-1 ms each and
// cudaFree synchronises the device; one fb_phase_blocks call makes ~30 temporary allocations.  Freed blocks are kept and
// handed back best-fit to the SAME context only: a context launches everything on its one stream, so reuse of a block is
// stream-ordered.  (A block never migrates to another context / stream, so a second context on the same device cannot be
// handed memory that the first one's kernels are still using.)  A process-wide registry maps live pointers to their
// owner so that fb_cache_free() needs no context argument.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <unordered_map>

struct FbCache;
struct FbCacheRegistry {
    std::mutex mu;
    std::unordered_map<void *, std::pair<FbCache *, size_t>> live;  // ptr -> (owner, size)
    static FbCacheRegistry &get() {
        static FbCacheRegistry r;
        return r;
    }
    // the owner is going away: blocks it handed out and that are still live are plain cudaMalloc memory from now on
    void orphan(FbCache *c) {
        std::lock_guard<std::mutex> g(mu);
        for (auto it = live.begin(); it != live.end();)
            if (it->second.first == c)
                it = live.erase(it);
            else
                ++it;
    }
};

struct FbCache {
    std::mutex mu;
    size_t live_bytes = 0;  // handed out and not yet released (cudaMemGetInfo costs tens of ms on a busy context: callers
                            // budget against the device total minus this figure instead)
    std::multimap<size_t, void *> free_blocks;
    cudaError_t alloc(void **p, size_t bytes) {
        if (bytes == 0) bytes = 1;
        bytes = (bytes + 511) & ~(size_t)511;
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = free_blocks.lower_bound(bytes);
            if (it != free_blocks.end() && it->first <= bytes + bytes / 2 + (1 << 20)) {  // best fit, bounded waste
                *p = it->second;
                const size_t sz = it->first;
                free_blocks.erase(it);
                live_bytes += sz;
                std::lock_guard<std::mutex> g2(FbCacheRegistry::get().mu);
                FbCacheRegistry::get().live[*p] = std::make_pair(this, sz);
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) {
            trim();  // give cached memory back and retry once
            cudaGetLastError();
            e = cudaMalloc(p, bytes);
            if (e != cudaSuccess) return e;
        }
        {
            std::lock_guard<std::mutex> g(mu);
            live_bytes += bytes;
        }
        std::lock_guard<std::mutex> g2(FbCacheRegistry::get().mu);
        FbCacheRegistry::get().live[*p] = std::make_pair(this, bytes);
        return cudaSuccess;
    }
    // the calling thread's current device must be the context's device
    void trim() {
        std::lock_guard<std::mutex> g(mu);
        for (auto &kv : free_blocks) cudaFree(kv.second);
        free_blocks.clear();
    }
};

static inline void fb_cache_free(void *p) {
    if (!p) return;
    FbCache *owner = nullptr;
    size_t sz = 0;
    {
        FbCacheRegistry &r = FbCacheRegistry::get();
        std::lock_guard<std::mutex> g(r.mu);
        auto it = r.live.find(p);
        if (it != r.live.end()) {
            owner = it->second.first;
            sz = it->second.second;
            r.live.erase(it);
        }
    }
    if (!owner) {
        cudaFree(p);  // not ours
        return;
    }
    std::lock_guard<std::mutex> g(owner->mu);
    owner->free_blocks.insert(std::make_pair(sz, p));
    owner->live_bytes -= sz < owner->live_bytes ? sz : owner->live_bytes;
}
