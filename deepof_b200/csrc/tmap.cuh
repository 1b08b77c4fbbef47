// TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point; no -lcuda) shared by the kernels that
// stage or drain activation tiles with cp.async.bulk.tensor.
#pragma once
#include <cuda.h>
#include <vector>
#include "common.cuh"

typedef CUresult (*dof_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static dof_tmap_encode_fn tmap_encoder() {
    static dof_tmap_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<dof_tmap_encode_fn>(p);
        cudaGetLastError();
    }
    return fn;
}


// 3-D TMA view of a [S, T, 2H] fp32 activation: box = [H columns, 1 step, R sequences], hardware swizzle of the row width
struct Tmap3Key { const void* base; int S, T, H, R; };
struct Tmap3Entry { Tmap3Key k; CUtensorMap m; };
static std::vector<Tmap3Entry> g_tmaps3;

static int tmap_seq3d(const float* base, int S, int T, int H, int R, CUtensorMap* out) {
    for (const Tmap3Entry& e : g_tmaps3)
        if (e.k.base == base && e.k.S == S && e.k.T == T && e.k.H == H && e.k.R == R) { *out = e.m; return DOF_OK; }
    dof_tmap_encode_fn enc = tmap_encoder();
    if (!enc) DOF_FAIL(DOF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    Tmap3Entry e;
    e.k = Tmap3Key{base, S, T, H, R};
    memset(&e.m, 0, sizeof(e.m));
    cuuint64_t dims[3] = {(cuuint64_t)(2 * H), (cuuint64_t)T, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)2 * H * 4, (cuuint64_t)T * 2 * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)H, 1, (cuuint32_t)R};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&e.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, H == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DOF_FAIL(DOF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a [%d, %d, %d] view", (int)r, S, T, 2 * H);
    if (g_tmaps3.size() >= 64) g_tmaps3.clear();
    g_tmaps3.push_back(e);
    *out = e.m;
    return DOF_OK;
}

