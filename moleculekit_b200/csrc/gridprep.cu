// gridprep.cu -- SURVEY 8f row 1: the host numpy passes either side of K1, on the device.
//
//   mkb_grid_centers   voxel centres exactly as getCenters / _getGridCenters build them
//                      (moleculekit/tools/voxeldescriptors.py:116-123,243-247): fl(fl(i * voxelsize) + bb_min[d]) in float64,
//                      (M, 3) rows in (ix*ny + iy)*nz + iz order -- for consumers that want them resident; K1 never
//                      reads a centre array.
//   mkb_rotate_coords  rotateCoordinates (voxeldescriptors.py:78-114), batched: three successive rotations about a
//                      centre, new = (x - c) . R^T + c in float64 (the reference's numpy promotes the float32
//                      coordinates to float64 there), one molecule = one (3 matrices, centre) pair, so a batch of poses
//                      gets an independent random rotation each without leaving the GPU.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace mkb {

struct CentersGrid {
    double origin[3];
    double vs;
    int dims[3];
    long long out_offset;
};

__global__ void grid_centers_kernel(const CentersGrid *__restrict__ grids, long long max_vox, double *__restrict__ out) {
    const CentersGrid &g = grids[blockIdx.y];
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long M = (long long)g.dims[0] * g.dims[1] * g.dims[2];
    if (v >= M) return;
    const int iz = (int)(v % g.dims[2]);
    const long long xy = v / g.dims[2];
    const int iy = (int)(xy % g.dims[1]), ix = (int)(xy / g.dims[1]);
    double *o = out + (g.out_offset + v) * 3;
    o[0] = __dadd_rn(__dmul_rn((double)ix, g.vs), g.origin[0]);
    o[1] = __dadd_rn(__dmul_rn((double)iy, g.vs), g.origin[1]);
    o[2] = __dadd_rn(__dmul_rn((double)iz, g.vs), g.origin[2]);
}

__global__ void rotate_coords_kernel(const float *__restrict__ in, long long n, const long long *__restrict__ atom_offsets,
                                     int B, const double *__restrict__ mats, const double *__restrict__ centers,
                                     float *__restrict__ out32, double *__restrict__ out64) {
    const long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (a >= n) return;
    int lo = 0, hi = B;  // molecule b with atom_offsets[b] <= a < atom_offsets[b+1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (atom_offsets[mid] <= a) lo = mid; else hi = mid;
    }
    const double *R = mats + (long long)lo * 27, *c = centers + (long long)lo * 3;
    double x[3] = {(double)in[3 * a], (double)in[3 * a + 1], (double)in[3 * a + 2]};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double d0 = __dsub_rn(x[0], c[0]), d1 = __dsub_rn(x[1], c[1]), d2 = __dsub_rn(x[2], c[2]);
        double y[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double *row = R + r * 9 + j * 3;  // (d . R^T)[j] = sum_k d[k] * R[j][k], left to right, no FMA
            const double s = __dadd_rn(__dadd_rn(__dmul_rn(d0, row[0]), __dmul_rn(d1, row[1])), __dmul_rn(d2, row[2]));
            y[j] = __dadd_rn(s, c[j]);
        }
        x[0] = y[0]; x[1] = y[1]; x[2] = y[2];
    }
    if (out64) { out64[3 * a] = x[0]; out64[3 * a + 1] = x[1]; out64[3 * a + 2] = x[2]; }
    if (out32) { out32[3 * a] = (float)x[0]; out32[3 * a + 1] = (float)x[1]; out32[3 * a + 2] = (float)x[2]; }
}

}  // namespace mkb

using namespace mkb;

extern "C" int mkb_grid_centers(mkb_handle_t h, void *stream, const mkb_grid_desc *grids, int32_t B, double *centers) {
    MKB_ENTER(h);
    cudaStream_t st = (cudaStream_t)stream;
    MKB_STREAM_ORDER(h, st);
    if (B < 0) return fail(h, MKB_ERR_BAD_ARG, "negative size");
    if (B == 0) return MKB_OK;
    if (!grids || !centers) return fail(h, MKB_ERR_BAD_ARG, "null grids/centers");
    if (B > 65535) return fail(h, MKB_ERR_BAD_ARG, "at most 65535 grids per call");
    std::vector<CentersGrid> cg((size_t)B);
    long long maxv = 0;
    for (int b = 0; b < B; ++b) {
        long long m = 1;
        for (int d = 0; d < 3; ++d) {
            if (grids[b].dims[d] <= 0) return fail(h, MKB_ERR_BAD_ARG, "grid %d: dims must be positive", b);
            cg[b].origin[d] = grids[b].origin[d];
            cg[b].dims[d] = grids[b].dims[d];
            m *= grids[b].dims[d];
        }
        if (grids[b].out_offset < 0) return fail(h, MKB_ERR_BAD_ARG, "grid %d: negative out_offset", b);
        cg[b].vs = grids[b].voxelsize;
        cg[b].out_offset = grids[b].out_offset;
        maxv = std::max(maxv, m);
    }
    if (cdiv(maxv, 256) >= (1ll << 31)) return fail(h, MKB_ERR_BAD_ARG, "grid too large");
    CentersGrid *d_cg;
    int rc;
    if ((rc = scratch_get(h, S_DESC, (size_t)B, &d_cg))) return rc;
    MKB_CUDA(h, cudaMemcpyAsync(d_cg, cg.data(), sizeof(CentersGrid) * (size_t)B, cudaMemcpyHostToDevice, st));
    // (pageable source: the runtime has staged the bytes when cudaMemcpyAsync returns, as for K1's descriptors)
    grid_centers_kernel<<<dim3((unsigned)cdiv(maxv, 256), (unsigned)B), 256, 0, st>>>(d_cg, maxv, centers);
    MKB_LAUNCHED(h);
    return MKB_OK;
}

extern "C" int mkb_rotate_coords(mkb_handle_t h, void *stream, const float *coords, int64_t n_atoms,
                                 const int64_t *atom_offsets, int32_t B, const double *matrices, const double *centers,
                                 float *out_f32, double *out_f64) {
    MKB_ENTER(h);
    cudaStream_t st = (cudaStream_t)stream;
    MKB_STREAM_ORDER(h, st);
    if (B < 0 || n_atoms < 0) return fail(h, MKB_ERR_BAD_ARG, "negative size");
    if (n_atoms == 0 || B == 0) return MKB_OK;
    if (!coords || !atom_offsets || !matrices || !centers || (!out_f32 && !out_f64))
        return fail(h, MKB_ERR_BAD_ARG, "null argument");
    rotate_coords_kernel<<<(unsigned)cdiv(n_atoms, 256), 256, 0, st>>>(coords, n_atoms, (const long long *)atom_offsets, B,
                                                                       matrices, centers, out_f32, out_f64);
    MKB_LAUNCHED(h);
    return MKB_OK;
}
