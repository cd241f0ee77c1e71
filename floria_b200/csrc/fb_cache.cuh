// fb_cache.cuh — caching device allocator (per process, per device).  cudaMalloc/cudaFree cost ~0.1-1 ms each and
// cudaFree synchronises the device; one fb_phase_blocks call makes ~30 temporary allocations, and with one process per
// GPU the driver lock is shared by all ranks.  Freed blocks are kept and handed back best-fit; everything the library
// launches is on one stream per context, so reuse is stream-ordered.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <unordered_map>

struct FbCache {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks[16];       // per device
    std::unordered_map<void *, std::pair<int, size_t>> live;  // ptr -> (device, size)
    static FbCache &get() {
        static FbCache c;
        return c;
    }
    cudaError_t alloc(void **p, size_t bytes) {
        if (bytes == 0) bytes = 1;
        bytes = (bytes + 511) & ~(size_t)511;
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> g(mu);
            auto &fl = free_blocks[dev & 15];
            auto it = fl.lower_bound(bytes);
            if (it != fl.end() && it->first <= bytes + bytes / 2 + (1 << 20)) {  // best fit, bounded waste
                *p = it->second;
                live[*p] = std::make_pair(dev, it->first);
                fl.erase(it);
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) {
            trim(dev);  // give cached memory back and retry once
            cudaGetLastError();
            e = cudaMalloc(p, bytes);
            if (e != cudaSuccess) return e;
        }
        std::lock_guard<std::mutex> g(mu);
        live[*p] = std::make_pair(dev, bytes);
        return cudaSuccess;
    }
    void release(void *p) {
        if (!p) return;
        std::lock_guard<std::mutex> g(mu);
        auto it = live.find(p);
        if (it == live.end()) {
            cudaFree(p);  // not ours
            return;
        }
        free_blocks[it->second.first & 15].insert(std::make_pair(it->second.second, p));
        live.erase(it);
    }
    void trim(int dev) {
        std::lock_guard<std::mutex> g(mu);
        auto &fl = free_blocks[dev & 15];
        for (auto &kv : fl) cudaFree(kv.second);
        fl.clear();
    }
};

static inline void fb_cache_free(void *p) { FbCache::get().release(p); }
