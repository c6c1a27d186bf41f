// launch_cache.cuh -- launch geometry looked up once per (device, kernel, block size, dynamic shared memory): the SM count, the
// resident CTAs per SM, and the opt-in for more than 48 KB of dynamic shared memory.  A batch is ~40 launches; asking the runtime
// for these numbers on every one of them made a one-read call (what nuclseq_search_bwa issues, reference extension.cpp:311-343)
// host-bound at ~1 ms.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <map>
#include <mutex>
#include <tuple>

inline int cached_sm_count() {
    static std::map<int, int> m;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    auto it = m.find(dev);
    if (it != m.end()) return it->second;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    m[dev] = sms;
    return sms;
}

// resident CTAs per SM (>= 1); raises the kernel's dynamic shared-memory limit first when smem > 48 KB
template <class K> int cached_blocks_per_sm(K kernel, int threads, size_t smem) {
    static std::map<std::tuple<int, const void*, int, size_t>, int> m;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_tuple(dev, reinterpret_cast<const void*>(kernel), threads, smem);
    std::lock_guard<std::mutex> g(mu);
    auto it = m.find(key);
    if (it != m.end()) return it->second;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem);
    if (nb < 1) nb = 1;
    m[key] = nb;
    return nb;
}
