// Window loader: raw pose frames -> model-ready windows x[B,T,N,3], a[B,T,E,1].
//
// One kernel replaces the reference's CPU chain between a per-video pose table and its window
// store (paths relative to the reference root):
//   centre            deepof/data.py:1844-1869
//   align             deepof/data.py:1878-1928 -> deepof/utils.py:2097-2142, rotate utils.py:1298-1319
//   speed             deepof/utils.py:3788-3857   rolling_speed(window=3, shift=2, rounds=3)
//   edge length       deepof/utils.py:863-881
//   size-normalise, log1p, per-video + global scalers   utils.py:2425-2566, 2866-2921
//   clip |z|>10 -> NaN -> linear interpolation (both directions), fillna(0)   utils.py:2990-3004, 2577-2583
//   rolling windows   deepof/utils.py:3354-3377, layout deepof/clustering/dataset.py:16-26
//
// Every scaler of the chain is affine, so the host folds them into one (scale, shift) pair per
// column; the kernel sees only those.  Arithmetic is fp64 (the reference works on float64 tables and
// rounds speeds to 3 decimals, which fp32 cannot reproduce); the kernel is bound by the fp32 tile
// writes, not by the math.
//
// Data movement per CTA (WPB consecutive windows, which share all but `step` frames with their
// neighbours): ONE 1-D TMA bulk copy (cp.async.bulk, mbarrier complete_tx) brings the contiguous
// [(WPB-1)*step + T + 4, 2N] fp32 frame slab into shared memory; features are computed once per frame
// into a shared fp64 tile; the overlapping windows are expanded into a shared fp32 staging tile in the
// exact output order and leave with TWO TMA bulk stores (x and a), so HBM sees only full-line writes.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"   // smem_u32, mbarrier helpers, fence_async_smem

#define LD_MAXN 32
#define LD_MAXE 64
#define LD_THREADS 256

struct LoaderP {
    const float* frames;      // [n_frames, N, 2] raw (x, y) per body part, one video
    long long n_frames;
    int T, step, N, E, center_node, align_node;
    double cx, cy, fps, clip, coord_scale, coord_shift;
    double speed_scale[LD_MAXN], speed_shift[LD_MAXN];
    double dist_div[LD_MAXE], dist_scale[LD_MAXE], dist_shift[LD_MAXE];
    short e0[LD_MAXE], e1[LD_MAXE];
    long long w_start;        // first window (index inside this video)
    int B, wpb, bulk_in, bulk_out;
    float* x;                 // [B, T, N, 3]
    float* a;                 // [B, T, E, 1]
};

__device__ __forceinline__ void ld_centre(const float* R, long long fb, long long f, const LoaderP& p, double& ox, double& oy) {
    if (p.center_node >= 0) {
        const float* q = R + (f - fb) * 2 * p.N + 2 * p.center_node;
        ox = (double)q[0]; oy = (double)q[1];
    } else { ox = p.cx; oy = p.cy; }
}

// rotation that puts the align body part on +y: angle = atan2(x_align, y_align)  (utils.py:2125)
__device__ __forceinline__ void ld_rot(const float* R, long long fb, long long f, const LoaderP& p, double& cs, double& sn) {
    if (p.align_node < 0) { cs = 1.0; sn = 0.0; return; }
    double ox, oy;
    ld_centre(R, fb, f, p, ox, oy);
    const float* q = R + (f - fb) * 2 * p.N + 2 * p.align_node;
    double ang = atan2((double)q[0] - ox, (double)q[1] - oy);
    sincos(ang, &sn, &cs);
}

// standardised feature `col` of frame f (NaN = undefined or clipped).  Columns: [0,2N) coords (node-major,
// x then y), [2N,3N) speeds, [3N,3N+E) log1p edge lengths.  R/fb: frame f lives at R[(f-fb)*2N ...].
template <typename COLS>
__device__ __forceinline__ double ld_feat(const float* R, long long fb, long long f, int col, const LoaderP& p, const COLS& cc, double cs, double sn) {
    const int N = p.N;
    double z;
    if (col < 2 * N) {
        const int n = col >> 1, c = col & 1;
        double ox, oy;
        ld_centre(R, fb, f, p, ox, oy);
        const float* q = R + (f - fb) * 2 * N + 2 * n;
        const double vx = (double)q[0] - ox, vy = (double)q[1] - oy;
        double r = c == 0 ? cs * vx - sn * vy : sn * vx + cs * vy;        // utils.py:1313-1317
        if (p.align_node >= 0 && fabs(r) < 1e-5) r = 0.0;                 // data.py:1912
        z = r * p.coord_scale + p.coord_shift;
    } else if (col < 3 * N) {
        const int n = col - 2 * N;
        if (f < 4) return nan("");                                        // shift(2) + rolling(3)
        double s = 0.0;
#pragma unroll
        for (int k = 2; k >= 0; k--) {
            const float* qa = R + (f - k - fb) * 2 * N + 2 * n;
            const float* qb = qa - 4 * N;
            const double dx = (double)qa[0] / 2.0 - (double)qb[0] / 2.0, dy = (double)qa[1] / 2.0 - (double)qb[1] / 2.0;
            s += sqrt(dx * dx + dy * dy);                                 // utils.py:3826-3838
        }
        const double v = rint(s / 3.0 * 1000.0) / 1000.0 * p.fps;         // np.round(., 3) * frame_rate
        z = v * cc.speed_scale[n] + cc.speed_shift[n];
    } else {
        const int e = col - 3 * N;
        const float* qa = R + (f - fb) * 2 * N + 2 * cc.e0[e];
        const float* qb = R + (f - fb) * 2 * N + 2 * cc.e1[e];
        const double dx = (double)qa[0] - (double)qb[0], dy = (double)qa[1] - (double)qb[1];
        double d = sqrt(dx * dx + dy * dy) / cc.dist_div[e];               // utils.py:877-880, 2520-2526
        if (d < 0.0) d = 0.0;
        z = log1p(d) * cc.dist_scale[e] + cc.dist_shift[e];                 // utils.py:2528-2531
    }
    if (p.clip > 0.0 && fabs(z) > p.clip) z = nan("");                    // utils.py:2996-2999
    return z;
}

template <typename COLS>
__device__ __forceinline__ double ld_feat_global(long long f, int col, const LoaderP& p, const COLS& cc) {
    double cs = 1.0, sn = 0.0;
    if (col < 2 * p.N) ld_rot(p.frames, 0, f, p, cs, sn);
    return ld_feat(p.frames, 0, f, col, p, cc, cs, sn);
}

// per-column constants as the kernels index them (shared-memory copy: a lane-dependent index into the kernel
// parameter space would serialise on the constant cache)
struct LoaderCols {
    double speed_scale[LD_MAXN], speed_shift[LD_MAXN];
    double dist_div[LD_MAXE], dist_scale[LD_MAXE], dist_shift[LD_MAXE];
    short e0[LD_MAXE], e1[LD_MAXE];
};
struct LoaderSmem { size_t raw, z, valid, rot, out, total; };
static inline LoaderSmem loader_smem(int T, int step, int N, int E, int wpb) {
    const size_t nf = (size_t)(wpb - 1) * step + T, C = 3 * (size_t)N + E;
    LoaderSmem s;
    size_t off = 16;                                        // mbarrier
    s.raw = off; off += ((nf + 4) * 2 * N * 4 + 32 + 15) / 16 * 16;
    s.z = off; off += nf * C * 8;
    s.rot = off; off += nf * 2 * 8;
    s.out = off; off += ((size_t)wpb * T * C * 4 + 15) / 16 * 16;
    s.valid = off; off += (nf * C + 15) / 16 * 16;
    s.total = off;
    return s;
}

__global__ void __launch_bounds__(LD_THREADS) load_windows_kernel(const __grid_constant__ LoaderP p, const LoaderSmem so) {
    extern __shared__ __align__(128) unsigned char ld_smem[];
    __shared__ LoaderCols cc;
    const int N = p.N, E = p.E, T = p.T, C = 3 * N + E, tid = threadIdx.x;
    for (int i = tid; i < LD_MAXN; i += LD_THREADS) { cc.speed_scale[i] = p.speed_scale[i]; cc.speed_shift[i] = p.speed_shift[i]; }
    for (int i = tid; i < LD_MAXE; i += LD_THREADS) {
        cc.dist_div[i] = p.dist_div[i]; cc.dist_scale[i] = p.dist_scale[i]; cc.dist_shift[i] = p.dist_shift[i];
        cc.e0[i] = p.e0[i]; cc.e1[i] = p.e1[i];
    }
    const int wl0 = blockIdx.x * p.wpb;                     // first window of this CTA inside the batch
    const int nw = min(p.wpb, p.B - wl0);
    const long long f_lo = (p.w_start + wl0) * (long long)p.step;
    const int nfr = (nw - 1) * p.step + T;
    const long long f_raw = f_lo >= 4 ? f_lo - 4 : 0;

    double* Z = reinterpret_cast<double*>(ld_smem + so.z);
    double* rot = reinterpret_cast<double*>(ld_smem + so.rot);
    unsigned char* valid = ld_smem + so.valid;
    float* outx = reinterpret_cast<float*>(ld_smem + so.out);
    float* outa = outx + (size_t)nw * T * 3 * N;
    const uint32_t bar = smem_u32(ld_smem);

    // ---- phase 1: frame slab -> shared memory (TMA bulk copy of the 16-byte aligned part) -------
    const long long b0 = f_raw * 2 * N * 4, b1 = (f_lo + nfr) * 2 * N * 4;       // byte range in `frames`
    const long long a0 = b0 & ~15LL;
    long long a1 = (b1 + 15) & ~15LL;
    const long long tot = (p.n_frames * 2 * N * 4) & ~15LL;
    if (a1 > tot) a1 = tot;
    unsigned char* rawb = ld_smem + so.raw;
    const float* R = reinterpret_cast<const float*>(rawb + (b0 - a0));           // frame f_raw
    const bool bulk = p.bulk_in && a1 > a0;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (bulk) {
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(a1 - a0);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(rawb)), "l"(reinterpret_cast<const unsigned char*>(p.frames) + a0), "r"(bytes), "r"(bar)
                         : "memory");
        }
        // unaligned tail of the table (at most 12 bytes)
        for (long long o = (a1 > b0 ? a1 : b0) + 4LL * tid; o < b1; o += 4LL * LD_THREADS)
            *reinterpret_cast<float*>(rawb + (o - a0)) = *reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p.frames) + o);
        mbar_wait(bar, 0);
    } else {
        for (long long o = b0 + 4LL * tid; o < b1; o += 4LL * LD_THREADS)
            *reinterpret_cast<float*>(rawb + (o - a0)) = *reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p.frames) + o);
    }
    __syncthreads();

    // ---- phase 2: every feature of every frame once; one warp per frame, lanes over the columns ---------
    const int warp = tid >> 5, lane = tid & 31, nwarp = LD_THREADS / 32;
    for (int fr = warp; fr < nfr; fr += nwarp) {
        double cs = 1.0, sn = 0.0;
        if (p.align_node >= 0) {
            if (lane == 0) ld_rot(R, f_raw, f_lo + fr, p, cs, sn);
            cs = __shfl_sync(0xffffffffu, cs, 0);
            sn = __shfl_sync(0xffffffffu, sn, 0);
        }
        for (int col = lane; col < C; col += 32) {
            const double z = ld_feat(R, f_raw, f_lo + fr, col, p, cc, cs, sn);
            Z[fr * C + col] = z;
            valid[fr * C + col] = (z == z) ? 1 : 0;
        }
    }
    __syncthreads();

    // ---- phase 3: clipped / undefined entries -> linear interpolation between the nearest valid
    // frames of the same column (pandas interpolate(limit_direction="both"), then fillna(0)).
    // Anchors are searched in the tile first, then in the frame table itself.
    for (int fr = warp; fr < nfr; fr += nwarp) {
        for (int col = lane; col < C; col += 32) {
            const int i = fr * C + col;
            if (valid[i]) continue;
            long long fp = -1, fn = -1;
            double zp = 0.0, zn = 0.0;
            for (int g = fr - 1; g >= 0; g--)
                if (valid[g * C + col]) { fp = f_lo + g; zp = Z[g * C + col]; break; }
            if (fp < 0)
                for (long long f = f_lo - 1; f >= 0; f--) {
                    const double z = ld_feat_global(f, col, p, cc);
                    if (z == z) { fp = f; zp = z; break; }
                }
            for (int g = fr + 1; g < nfr; g++)
                if (valid[g * C + col]) { fn = f_lo + g; zn = Z[g * C + col]; break; }
            if (fn < 0)
                for (long long f = f_lo + nfr; f < p.n_frames; f++) {
                    const double z = ld_feat_global(f, col, p, cc);
                    if (z == z) { fn = f; zn = z; break; }
                }
            double v = 0.0;
            if (fp >= 0 && fn >= 0) v = zp + (zn - zp) / (double)(fn - fp) * (double)(f_lo + fr - fp);
            else if (fp >= 0) v = zp;
            else if (fn >= 0) v = zn;
            Z[i] = v;
        }
    }
    __syncthreads();

    // ---- phase 4: expand the overlapping windows in output order (one warp per (window, step) row, lanes over
    // the row's columns; x column q = 3n + c reads feature 2n + c or 2N + n), then TMA bulk stores -----------
    const int NX = T * 3 * N, NA = T * E;
    for (int row = warp; row < nw * T; row += nwarp) {
        const int w = row / T, t = row - w * T;
        const double* zr = Z + (size_t)(w * p.step + t) * C;
        float* ox = outx + (size_t)row * 3 * N;
        for (int q = lane; q < 3 * N; q += 32) {
            const int n = q / 3, c = q - 3 * n;
            ox[q] = (float)zr[c < 2 ? 2 * n + c : 2 * N + n];
        }
        float* oa = outa + (size_t)row * E;
        for (int e = lane; e < E; e += 32) oa[e] = (float)zr[3 * N + e];
    }
    float* gx = p.x + (size_t)wl0 * NX;
    float* ga = p.a + (size_t)wl0 * NA;
    const bool bulk_out = p.bulk_out && ((nw * NX) % 4 == 0) && ((nw * NA) % 4 == 0) && ((wl0 * (long long)NX) % 4 == 0) &&
                          ((wl0 * (long long)NA) % 4 == 0);
    if (bulk_out) {
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gx), "r"(smem_u32(outx)), "r"((uint32_t)(nw * NX * 4)) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(ga), "r"(smem_u32(outa)), "r"((uint32_t)(nw * NA * 4)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        __syncthreads();
        for (int o = tid; o < nw * NX; o += LD_THREADS) gx[o] = outx[o];
        for (int o = tid; o < nw * NA; o += LD_THREADS) ga[o] = outa[o];
    }
}

// Per-frame |p_a - p_b| (the size reference Nose - Tail_base, utils.py:2478-2489); the host takes the nanmedian.
__global__ void loader_pair_length_kernel(const float* __restrict__ frames, long long n_frames, int N, int na, int nb,
                                          double* __restrict__ out) {
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < n_frames; f += (long long)gridDim.x * blockDim.x) {
        const float* q = frames + f * 2 * N;
        out[f] = hypot((double)q[2 * na] - (double)q[2 * nb], (double)q[2 * na + 1] - (double)q[2 * nb + 1]);
    }
}

// Moments of the standardised columns of one video, NaNs skipped, per column group g in {coords, speeds,
// edge lengths}: out[3g+0] += count, out[3g+1] += sum(z - shift[g]), out[3g+2] += sum((z - shift[g])^2).
// This is what the groupwise StandardScaler fits see (utils.py:2547-2563, 2665-2793) when the host sets the
// per-column (scale, shift) of `p` to the stage whose statistics it wants.
__global__ void __launch_bounds__(256) loader_moments_kernel(const __grid_constant__ LoaderP p, double s0, double s1, double s2,
                                                             double* __restrict__ out) {
    const int N = p.N, C = 3 * N + p.E;
    double acc[9];
#pragma unroll
    for (int i = 0; i < 9; i++) acc[i] = 0.0;
    const long long total = p.n_frames * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / C;
        const int col = (int)(i - f * C);
        const double z = ld_feat_global(f, col, p, p);
        if (z != z) continue;
        const int g = col < 2 * N ? 0 : (col < 3 * N ? 1 : 2);
        const double d = z - (g == 0 ? s0 : (g == 1 ? s1 : s2));
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (g == k) { acc[3 * k] += 1.0; acc[3 * k + 1] += d; acc[3 * k + 2] += d * d; }
    }
    __shared__ double red[9][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const double v = warp_sum_d(acc[i]);
        if (lane == 0) red[i][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, v);
    }
}
