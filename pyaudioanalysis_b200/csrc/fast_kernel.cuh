// Register-tiled short-term kernels specialised per window length.
#pragma once
#include <vector>
#include "common.cuh"

namespace b200aa {

struct FastTables {
    void release() {}
};

inline int fast_plan_init(int fs, int window, int step, const std::vector<int> &blob, const BlobLayout &bl,
                          FastTables *ft, int *kind)
{
    (void)fs; (void)window; (void)step; (void)blob; (void)bl; (void)ft;
    *kind = 0;
    return B200AA_OK;
}

inline int fast_launch_features(int kind, const FastTables &ft, const StParams &p, int sm_count, int64_t T, cudaStream_t st)
{
    (void)kind; (void)ft; (void)p; (void)sm_count; (void)T; (void)st;
    return B200AA_ERR_UNSUPPORTED;
}

}  // namespace b200aa
