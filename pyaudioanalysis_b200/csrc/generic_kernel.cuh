// Generic short-term kernel: any window length.  Mixed-radix Stockham transform in shared
// memory (one output element per thread per pass, arbitrary prime radices), then one warp per
// frame for the features.  This is the correctness baseline for every (fs, window, step); the
// register-tiled kernels in fast_kernel.cuh take over for the window lengths they specialise.
#pragma once
#include "common.cuh"
#include "dft_codelets.cuh"

namespace b200aa {

// bytes of the per-CTA arrays that scale with the window: transform ping-pong buffers + magnitude rows (+ previous frame)
inline size_t generic_big_bytes(int G, int Nc, int Kp)
{
    return (size_t(G) * Nc * sizeof(float2) * 2 + size_t(G + 1) * Kp * sizeof(float) + 15) & ~size_t(15);
}
// shared-memory bytes of the generic kernel for G frames per group (big = false: the window-sized arrays live in global memory)
inline size_t generic_smem_bytes(int G, int Nc, int Kp, int blob_words, bool big_in_smem = true)
{
    size_t b = big_in_smem ? generic_big_bytes(G, Nc, Kp) : 0;
    b += size_t(G + 1) * kFvStride * sizeof(float);          // feature rows (+ previous frame)
    b += size_t(kWarps) * B200AA_N_MEL * sizeof(float);      // mel scratch
    b += size_t(G + 1) * sizeof(float);                      // row sums
    b += size_t(kWarps) * 64 * sizeof(float) + 32 * sizeof(int4) + 16;   // time-domain parts + lane constants
    b += size_t(blob_words) * sizeof(int);
    return (b + 15) & ~size_t(15);
}

// One Stockham pass of radix R with one BUTTERFLY per thread (R = 2, 3, 4, 5, 7: register codelets of
// dft_codelets.cuh); other radices use the one-output-per-thread form inside the kernel.
template <int R>
__device__ __forceinline__ void stockham_pass_bfly(const float2 *src, float2 *dst, int ng, int Nc, int Ns,
                                                   const float2 *__restrict__ tw, int tid)
{
    const int nb = Nc / R, tstep = Nc / (Ns * R);
    for (int e = tid; e < ng * nb; e += kThreads) {
        const int f = e / nb, j = e - f * nb;
        const int k = j % Ns;
        const float2 *in = src + size_t(f) * Nc + j;
        float2 v[R];
        v[0] = in[0];
        int idx = 0;
#pragma unroll
        for (int r = 1; r < R; ++r) {
            idx += k * tstep;                      // (k * r * tstep) < Nc because k < Ns and r < R
            v[r] = cmul(in[r * nb], __ldg(tw + idx));
        }
        dft_small<R>(v);
        float2 *out = dst + size_t(f) * Nc + (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) out[q * Ns] = v[q];
    }
}

// BIG: windows whose transform does not fit shared memory (e.g. the 1 s windows of music_thumbnailing at 44.1 kHz,
// audioSegmentation.py:1137-1139): same code, the window-sized arrays of the CTA sit in global memory (L2 resident),
// __syncthreads() orders the passes as before.
template <int MODE, bool BIG = false>
__global__ void __launch_bounds__(kThreads, 2) st_generic_kernel(const StParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = p.G, Nc = p.Nc, K = p.K, Kp = p.Kp, w = p.window, fn = p.fft_n;
    unsigned char *const big = BIG ? p.scratch + size_t(blockIdx.x) * p.scratch_stride : smem_raw;
    float2 *bufA = reinterpret_cast<float2 *>(big);
    float2 *bufB = bufA + size_t(G) * Nc;
    float *Xrows = reinterpret_cast<float *>(bufB + size_t(G) * Nc);    // row 0 = previous frame
    float *fvrows = BIG ? reinterpret_cast<float *>(smem_raw) : Xrows + size_t(G + 1) * Kp;   // row 0 = previous frame
    float *mscr = fvrows + size_t(G + 1) * kFvStride;
    float *rowsum = mscr + kWarps * B200AA_N_MEL;
    float *tparts = rowsum + ((G + 1 + 3) & ~3);
    int4 *tlane = reinterpret_cast<int4 *>(tparts + kWarps * 64);
    int *blob_s = reinterpret_cast<int *>(tlane + 32);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < p.bl.words; i += kThreads) blob_s[i] = p.blob[i];
    if (tid < 32) tlane[tid] = time_lane_init(p.fft_n, tid);
    __syncthreads();
    const SmallTables tb = bind_tables(blob_s, p.bl);

    for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int64_t b = item / p.segs_per_clip, seg = item % p.segs_per_clip;
        const int64_t len = p.len ? p.len[b] : p.n_samples;
        int64_t n_rows, n_valid, origin;
        if (MODE == kModeFeatures) {
            n_rows = n_valid = len < w ? 0 : (len - w) / p.step + 1;     // loop guard :608
            origin = 0;
        } else {
            n_rows = p.rows_launch;
            n_valid = p.rows_valid;
            origin = p.origin;
        }
        const int64_t t0 = seg * p.seg_len;
        if (t0 >= n_rows) continue;
        const int64_t t1 = min(t0 + p.seg_len, n_rows);
        const b200aa_clip_norm nm = p.norm[b];
        const char *clip = reinterpret_cast<const char *>(p.sig) +
                           size_t(b) * p.clip_stride * (p.dtype == B200AA_DTYPE_I16 ? 2 : 4);
        const SampleReader rd{clip, p.dtype, nm.m};
        const int halo = (MODE == kModeFeatures) ? int(t0 < 2 ? t0 : 2) : 0;

        for (int64_t g0 = t0 - halo; g0 < t1; g0 += G) {
            const int ng = int((t1 - g0) < G ? (t1 - g0) : G);
            // ---- load: z[n] = (x[2n]-x0) + i (x[2n+1]-x0)   (or real only for odd windows)
            // Subtracting the frame's first sample makes constant frames transform to exact zeros,
            // as they (up to 1e-17) do in the float64 reference.
            for (int e = tid; e < ng * Nc; e += kThreads) {
                const int f = e / Nc, n = e - f * Nc;
                const int64_t fr = g0 + f;
                float2 z = make_float2(0.f, 0.f);
                if (fr < n_valid) {
                    const int64_t s0 = origin + fr * p.step;
                    const float d0 = rd(s0);
                    if (p.packed) z = make_float2(rd(s0 + 2 * n) - d0, rd(s0 + 2 * n + 1) - d0);
                    else z = make_float2(rd(s0 + n) - d0, 0.f);
                }
                bufA[e] = z;
            }
            __syncthreads();
            if (MODE == kModeFeatures) {
                // time-domain rows from the staged samples (bufA holds x - x[frame start]; one warp per frame).
                // fvrows rows of this step are free: the previous step's store loop finished before its last barrier.
                for (int f = warp; f < ng; f += kWarps) {
                    const float2 *zf = bufA + size_t(f) * Nc;
                    const float d0 = rd(origin + (g0 + f) * p.step);
                    float *fv = fvrows + size_t(f + 1) * kFvStride;
                    if (p.packed)
                        time_features_chunked([&](int n) { const float2 q = zf[n >> 1]; return ((n & 1) ? q.y : q.x) + d0; }, w, nm,
                                              tlane[lane], tparts + warp * 64, fv, lane);
                    else
                        time_features_chunked([&](int n) { return zf[n].x + d0; }, w, nm, tlane[lane], tparts + warp * 64, fv, lane);
                }
                // no barrier needed before the passes: they only read bufA as well
            }
            // ---- Stockham passes
            float2 *src = bufA, *dst = bufB;
            int Ns = 1;
            for (int ps = 0; ps < p.nrad; ++ps) {
                const int R = p.radix[ps];
                const int NsR = Ns * R, stride = Nc / R, tstep = Nc / NsR;
                if (R == 4) stockham_pass_bfly<4>(src, dst, ng, Nc, Ns, p.tw, tid);
                else if (R == 2) stockham_pass_bfly<2>(src, dst, ng, Nc, Ns, p.tw, tid);
                else if (R == 3) stockham_pass_bfly<3>(src, dst, ng, Nc, Ns, p.tw, tid);
                else if (R == 5) stockham_pass_bfly<5>(src, dst, ng, Nc, Ns, p.tw, tid);
                else if (R == 7) stockham_pass_bfly<7>(src, dst, ng, Nc, Ns, p.tw, tid);
                else
                for (int e = tid; e < ng * Nc; e += kThreads) {
                    const int f = e / Nc, o = e - f * Nc;
                    const int hi_ = o / NsR, rem = o - hi_ * NsR;
                    const int q = rem / Ns, k = rem - q * Ns;
                    const int j = hi_ * Ns + k;
                    int ph = (k * tstep + q * stride) % Nc;   // phase step per input
                    int idx = 0;
                    const float2 *in = src + size_t(f) * Nc + j;
                    float2 acc = make_float2(0.f, 0.f);
                    for (int r = 0; r < R; ++r) {
                        const float2 t = __ldg(p.tw + idx);
                        const float2 v = in[r * stride];
                        acc.x = fmaf(v.x, t.x, fmaf(-v.y, t.y, acc.x));
                        acc.y = fmaf(v.x, t.y, fmaf(v.y, t.x, acc.y));
                        idx += ph;
                        if (idx >= Nc) idx -= Nc;
                    }
                    dst[e] = acc;
                }
                __syncthreads();
                float2 *t_ = src; src = dst; dst = t_;
                Ns = NsR;
            }
            // ---- magnitudes |X[k]| / K, k < K  (ShortTermFeatures.py:617-621)
            for (int e = tid; e < ng * K; e += kThreads) {
                const int f = e / K, k = e - f * K;
                const int64_t fr = g0 + f;
                float mag = 0.f;
                if (fr < n_valid) {
                    const float2 *Z = src + size_t(f) * Nc;
                    const float sc = nm.a / float(K);
                    float2 Xc;
                    if (p.packed) {
                        // bins above fn/2 (only for a clipped frame with K > fn/2) mirror: |X[k]| = |X[fn-k]|
                        const int kk = k > Nc ? fn - k : k;
                        if (kk == Nc) {
                            Xc = make_float2(Z[0].x - Z[0].y, 0.f);          // Nyquist bin
                        } else {
                            const float2 zk = Z[kk];
                            const float2 zm = Z[kk == 0 ? 0 : Nc - kk];
                            const float2 ev = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                            const float2 od = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
                            const float2 wk = __ldg(p.tw_post + kk);
                            const float2 t = cmul(od, wk);
                            Xc = make_float2(ev.x + t.x, ev.y + t.y);
                        }
                    } else {
                        Xc = Z[k];
                    }
                    if (k == 0) {
                        // DC of y = a*(d - d0) + (a*d0 + bp):  a*sum(d-d0) + w*(a*d0+bp)
                        const float d0 = rd(origin + fr * p.step);
                        mag = fabsf(fmaf(nm.a, Xc.x, float(fn) * fmaf(nm.a, d0, nm.bp))) / float(K);
                    } else {
                        const float re = Xc.x * sc, im = Xc.y * sc;
                        mag = sqrtf(fmaf(re, re, im * im));
                    }
                }
                Xrows[size_t(f + 1) * Kp + k] = mag;
            }
            __syncthreads();

            if (MODE == kModeSpectrogram) {
                for (int e = tid; e < ng * K; e += kThreads) {
                    const int f = e / K, k = e - f * K;
                    p.out[(size_t(b) * p.rows_total + p.row0 + (g0 + f)) * K + k] = Xrows[size_t(f + 1) * Kp + k];
                }
                __syncthreads();
                continue;
            }
            if (MODE == kModeChromagram) {
                for (int f = warp; f < ng; f += kWarps) {
                    const float *X = Xrows + size_t(f + 1) * Kp;
                    float sxx = 0.f;
                    for (int k = lane; k < K; k += 32) sxx = fmaf(X[k], X[k], sxx);
                    sxx = warp_sum(sxx);
                    const float ch = (g0 + f < n_valid) ? chroma_lane(X, sxx, tb, lane) : 0.f;
                    if (lane < 12) p.out[(size_t(b) * p.rows_total + p.row0 + (g0 + f)) * 12 + lane] = ch;
                }
                __syncthreads();
                continue;
            }

            // ---- features: one warp per frame
            for (int f = warp; f < ng; f += kWarps) {
                const int64_t fr = g0 + f;
                const float *X = Xrows + size_t(f + 1) * Kp;
                // previous spectrum: row f (row 0 carries the last frame of the previous group);
                // the very first frame of a clip -- and a halo frame without history -- uses itself
                const bool has_prev = (fr > 0) && !(f == 0 && g0 == t0 - halo);
                float sxp;
                const float *Xp;
                float *fv = fvrows + size_t(f + 1) * kFvStride;
                if (has_prev && f > 0) {
                    // row sum of the neighbour is produced by another warp in this same phase:
                    // recompute it here instead of synchronising
                    Xp = Xrows + size_t(f) * Kp;
                    float s = 0.f;
                    for (int k = lane; k < K; k += 32) s += Xp[k];
                    sxp = warp_sum(s);
                } else if (has_prev) {
                    Xp = Xrows;
                    sxp = rowsum[0];
                } else {
                    Xp = X;
                    float s = 0.f;
                    for (int k = lane; k < K; k += 32) s += X[k];
                    sxp = warp_sum(s);
                }
                spectral_features(X, Xp, sxp, K, tb, mscr + warp * B200AA_N_MEL, fv, lane, rowsum + f + 1);
            }
            __syncthreads();
            // ---- store [n_out x ng] tile: consecutive threads -> consecutive frames
            for (int e = tid; e < p.n_out * ng; e += kThreads) {
                const int f = e / ng, c = e - f * ng;
                const int64_t fr = g0 + c;
                if (fr < t0) continue;
                float v;
                if (f < B200AA_N_BASE) v = fvrows[size_t(c + 1) * kFvStride + f];
                else {
                    const int fb = f - B200AA_N_BASE;
                    v = fr == 0 ? 0.f : fvrows[size_t(c + 1) * kFvStride + fb] - fvrows[size_t(c) * kFvStride + fb];
                }
                p.out[(size_t(b) * p.n_out + f) * p.t_stride + fr] = v;
            }
            __syncthreads();
            // ---- carry the last frame of the group into row 0
            for (int k = tid; k < K; k += kThreads) Xrows[k] = Xrows[size_t(ng) * Kp + k];
            if (tid < kFvStride) fvrows[tid] = fvrows[size_t(ng) * kFvStride + tid];
            if (tid == 0) rowsum[0] = rowsum[ng];
            __syncthreads();
        }
    }
}

}  // namespace b200aa
