// wrapping.cu -- K9: orthorhombic periodic wrapping of bonded groups for sm_100a.
//
// Replaces wrap_box (moleculekit/wrapping/wrapping.pyx:91-144), the loop Molecule.wrap runs for rectangular cells
// (moleculekit/molecule.py:2077).  Per frame the reference (1) takes the box centre as the running mean of the
// `centersel` atoms (pyx:113-118) or a fixed `center` (pyx:106-108), (2) for every bonded group [groups[g],
// groups[g+1]) takes the running mean of its atoms (pyx:127-134) and (3) per axis, when that centre is further than
// half a box from the box centre, subtracts box*round(diff/box) from every atom of the group (pyx:137-142).
//
// Frames, groups and axes are independent (groups are disjoint atom ranges and the box centre is read before any
// group of its frame is moved), so the parallel form is exact.  The running mean itself, c += (x - c)/(n + 1) with one
// float rounding per operation, is a sequential chain by definition and is kept verbatim (bit-identical output).
// The trajectory is frame-minor, so 32 consecutive frames of one (atom, axis) row are one 128-byte line: lanes = frames.
//
//   small groups (<= WRAP_SMALL = 4 atoms: waters, ions)  one thread per (group, frame): the 3*count values are loaded
//       once into registers, the three short chains run on them, moved axes are stored back from the registers.
//       Traffic: one read, one write of what moved -- the HBM-bound bulk of a solvated system.
//   long groups (a solute, lipids) and the centre selection  one CTA per (group, 32 frames): warps 3..7 stream the
//       rows through a 3-stage cp.async ring in shared memory while warps 0..2 (one axis each) walk the chain out of
//       shared memory with the reciprocal of n + 1 precomputed off the chain (exact_div.cuh); the chain (~6 dependent
//       float ops per atom), not DRAM latency, sets the pace.  The translation then goes to all 8 warps, which apply
//       it to the group's rows with 12 independent loads in flight per thread.
//   A classify pre-pass appends the long groups to a device list; the first CTAs of the groups launch are persistent
//   workers over that list, the rest are the small-group threads, so both kinds overlap on the machine.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "exact_div.cuh"

namespace mkb {

constexpr int WRAP_SMALL = 4;     // largest group handled from registers (waters, ions, 4-site water models)
constexpr int WRAP_STAGE = 32;    // atoms per pipeline stage of the long-chain path
constexpr int WRAP_NSTAGE = 3;    // cp.async ring depth
constexpr int WRAP_THREADS = 256;
constexpr int WRAP_WARPS = WRAP_THREADS / 32;
constexpr int WRAP_ROWS = WRAP_STAGE * 3;

struct alignas(16) ChainSmem {
    float v[WRAP_NSTAGE][WRAP_ROWS][32];  // [stage][atom*3 + axis][frame lane]
    float rb[WRAP_NSTAGE][WRAP_STAGE];    // (float)(n + 1)
    float rr[WRAP_NSTAGE][WRAP_STAGE];    // refined reciprocal of it
    float tr[3][32];                      // translation per (axis, frame lane)
    int mv[3][32];                        // the group moves along this axis in this frame
};

__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// 16-byte copy of which only the first src_bytes are read (the rest of the destination is zero-filled)
__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Running mean over `count` atoms for 32 frames [f0, f0 + nf) and the three axes, by the whole CTA.  Atom n of the chain
// is index[n] (INDIRECT: centersel, pyx:113-118) or first_atom + n (a group, pyx:127-134).  Warps 3..7 are the loaders
// (cp.async rows of 32 frames into the ring, plus the (n + 1, reciprocal) tables), warps 0..2 walk the chain of axis =
// warp and return its mean for frame lane (other warps return 0).
template <bool INDIRECT>
__device__ __forceinline__ float chain_32frames(ChainSmem &sm, const float *coords, long long fs, long long f0, int nf,
                                                const unsigned *__restrict__ index, long long first_atom,
                                                long long count) {
    constexpr int LOADERS = WRAP_WARPS - 3;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long long nst = (count + WRAP_STAGE - 1) / WRAP_STAGE;
    const float *lane_base = coords + f0 + lane;
    // rows start on 16-byte boundaries: 16-byte cp.async (4-byte copies are issued per element and ~4x slower)
    const bool aligned16 = ((reinterpret_cast<unsigned long long>(coords) | (unsigned long long)(fs * 4)) & 15ull) == 0;
    auto issue = [&](long long s) {
        if (w >= 3 && s < nst) {
            const int buf = (int)(s % WRAP_NSTAGE), lw = w - 3;
            const long long n0 = s * WRAP_STAGE;
            const long long left = count - n0;
            const int rows = left < WRAP_STAGE ? 3 * (int)left : WRAP_ROWS;  // row = atom_in_stage*3 + axis
            if (aligned16) {
                // 16-byte pieces: 8 per row; loader thread lt takes pieces lt, lt + 160, ...
                const int lt = t - 96;
                const float *p = coords + f0 + (INDIRECT ? 0 : (first_atom + n0) * 3 * fs);
#pragma unroll 5
                for (int piece = lt; piece < rows * 8; piece += LOADERS * 32) {
                    const int row = piece >> 3, q = piece & 7;
                    const int vf = nf - 4 * q;  // valid frames in this piece
                    if (vf > 0) {
                        long long off;
                        if (INDIRECT) {
                            const int a = row / 3, i = row - 3 * a;
                            off = ((long long)__ldg(index + n0 + a) * 3 + i) * fs;
                        } else {
                            off = row * fs;
                        }
                        cp_async16(&sm.v[buf][row][4 * q], p + off + 4 * q, vf >= 4 ? 16 : 4 * vf);
                    }
                }
            } else if (lane < nf) {
                if (INDIRECT) {
#pragma unroll 4
                    for (int row = lw; row < rows; row += LOADERS) {
                        const int a = row / 3, i = row - 3 * a;
                        cp_async4(&sm.v[buf][row][lane], lane_base + ((long long)__ldg(index + n0 + a) * 3 + i) * fs);
                    }
                } else {
                    const float *p = lane_base + (first_atom + n0) * 3 * fs;  // row r of the stage is p + r*fs
#pragma unroll 4
                    for (int row = lw; row < rows; row += LOADERS) cp_async4(&sm.v[buf][row][lane], p + row * fs);
                }
            }
            if (lw == 0) {
                const float b = __int2float_rn((int)(n0 + lane) + 1);  // the reference's (n + 1): C int -> float
                sm.rb[buf][lane] = b;
                sm.rr[buf][lane] = refined_rcp(b);
            }
        }
        cp_async_commit();  // one group per stage, empty past the end: keeps the wait distance uniform
    };
#pragma unroll
    for (int s = 0; s < WRAP_NSTAGE - 1; ++s) issue(s);
    float c = 0.f;
    for (long long s = 0; s < nst; ++s) {
        cp_async_wait<WRAP_NSTAGE - 2>();  // this thread's copies of stage s have landed (only stage s + 1 may be pending)
        __syncthreads();                   // ... and everyone's; the chain warps are also done with stage s - 1,
        issue(s + WRAP_NSTAGE - 1);        // whose buffer the loaders refill while the chain walks stage s
        if (w < 3) {
            const int buf = (int)(s % WRAP_NSTAGE);
            const long long left = count - s * WRAP_STAGE;
            const int m = left < WRAP_STAGE ? (int)left : WRAP_STAGE;
            // 8 atoms at a time: operands first (independent shared loads), then the dependent chain on the branch-free
            // fast division.  The range test runs beside the chain (OR of independent terms); when any step of any lane
            // left the fast domain (zero / denormal / huge difference) the batch is redone with __fdiv_rn.
            auto batch = [&](int a0, int nb) {  // nb = 8 (full, branch-free) or the tail length
                float x[8], b[8], r[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {  // a0 + u < WRAP_STAGE always: stale rows are read but never used
                    x[u] = sm.v[buf][(a0 + u) * 3 + w][lane];
                    b[u] = sm.rb[buf][a0 + u];
                    r[u] = sm.rr[buf][a0 + u];
                }
                const float c_in = c;
                unsigned bad = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (u < nb) {
                        const float d = __fsub_rn(x[u], c);
                        bad |= div_fast_ok(d) ? 0u : 1u;
                        c = __fadd_rn(c, div_fast(d, b[u], r[u]));
                    }
                if (bad) {
                    c = c_in;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (u < nb) c = __fadd_rn(c, __fdiv_rn(__fsub_rn(x[u], c), b[u]));
                }
            };
            int a0 = 0;
            for (; a0 + 8 <= m; a0 += 8) batch(a0, 8);
            if (a0 < m) batch(a0, m - a0);
        }
    }
    __syncthreads();  // the last stage is consumed before the caller reuses the ring
    cp_async_wait<0>();
    return c;
}

// box centre per (axis, frame): pyx:113-118.  One CTA per 32 frames.
__global__ void __launch_bounds__(WRAP_THREADS)
wrap_center_kernel(const float *__restrict__ coords, long long F, long long fs, const unsigned *__restrict__ centersel,
                   long long n_centersel, float *__restrict__ centre) {
    __shared__ ChainSmem sm;
    const long long f0 = 32ll * blockIdx.x;
    const int nf = (int)(F - f0 < 32 ? F - f0 : 32);
    const float c = chain_32frames<true>(sm, coords, fs, f0, nf, centersel, 0, n_centersel);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w < 3 && lane < nf) centre[(long long)w * F + f0 + lane] = c;
}

__global__ void wrap_classify_kernel(const unsigned *__restrict__ groups, long long n_ranges,
                                     unsigned *__restrict__ long_list, unsigned *__restrict__ n_long) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= n_ranges) return;
    if ((long long)groups[g + 1] - (long long)groups[g] > WRAP_SMALL) long_list[atomicAdd(n_long, 1u)] = (unsigned)g;
}

struct WrapArgs {
    float *coords;
    const float *box;
    long long F, fs, fsb;
    const unsigned *groups;
    long long n_ranges;
    const float *centre;  // [3][F], unused when fixed
    float cx, cy, cz;
    int fixed_centre;
    const unsigned *long_list;
    const unsigned *n_long;
    int n_long_ctas;
};

// pyx:137-139: does the group move along this axis, and by how much
__device__ __forceinline__ bool wrap_decide(const WrapArgs &A, float c, int i, long long f, float *tr) {
    const float bc = A.fixed_centre ? (i == 0 ? A.cx : (i == 1 ? A.cy : A.cz)) : A.centre[(long long)i * A.F + f];
    const float b = A.box[(long long)i * A.fsb + f];
    const float diff = __fsub_rn(c, bc);
    const bool move = fabsf(diff) > __fdiv_rn(b, 2.f);
    *tr = move ? __fmul_rn(b, roundf(__fdiv_rn(diff, b))) : 0.f;
    return move;
}

__device__ __forceinline__ void wrap_small_groups(const WrapArgs &A, long long tid);

// The first n_long_ctas CTAs are persistent workers over (long group, 32-frame chunk) items; the others are the
// small-group threads.  One launch, so the latency-bound long chains run beside the bandwidth-bound small groups.
__global__ void __launch_bounds__(WRAP_THREADS, 4) wrap_groups_kernel(const WrapArgs A) {
    __shared__ ChainSmem sm;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if ((int)blockIdx.x >= A.n_long_ctas) {
        wrap_small_groups(A, ((long long)blockIdx.x - A.n_long_ctas) * WRAP_THREADS + t);
        return;
    }
    {
        const long long nchunks = (A.F + 31) / 32;
        const long long total = (long long)(*A.n_long) * nchunks;
        for (long long item = blockIdx.x; item < total; item += A.n_long_ctas) {
            const long long slot = item / nchunks, f0 = 32 * (item - slot * nchunks);
            const int nf = (int)(A.F - f0 < 32 ? A.F - f0 : 32);
            const long long g = A.long_list[slot];
            const long long s = A.groups[g], count = (long long)A.groups[g + 1] - s;
            const float c = chain_32frames<false>(sm, A.coords, A.fs, f0, nf, nullptr, s, count);
            if (w < 3) {
                float tr = 0.f;
                const bool move = lane < nf && wrap_decide(A, c, w, f0 + lane, &tr);
                sm.tr[w][lane] = tr;
                sm.mv[w][lane] = move ? 1 : 0;
            }
            __syncthreads();
            const bool m0 = sm.mv[0][lane] != 0, m1 = sm.mv[1][lane] != 0, m2 = sm.mv[2][lane] != 0;
            const float t0 = sm.tr[0][lane], t1 = sm.tr[1][lane], t2 = sm.tr[2][lane];
            if (__syncthreads_or(m0 || m1 || m2)) {
                float *base = A.coords + s * 3 * A.fs + f0 + lane;
                constexpr int BATCH = 4;  // atoms per thread per round: 12 independent loads, then their stores
                for (long long k0 = w; k0 < count; k0 += WRAP_WARPS * BATCH) {
                    float v[BATCH][3];
#pragma unroll
                    for (int u = 0; u < BATCH; ++u) {
                        const long long k = k0 + WRAP_WARPS * u;
                        float *p = base + k * 3 * A.fs;
                        if (k < count) {
                            if (m0) v[u][0] = p[0];
                            if (m1) v[u][1] = p[A.fs];
                            if (m2) v[u][2] = p[2 * A.fs];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < BATCH; ++u) {
                        const long long k = k0 + WRAP_WARPS * u;
                        float *p = base + k * 3 * A.fs;
                        if (k < count) {
                            if (m0) p[0] = __fsub_rn(v[u][0], t0);
                            if (m1) p[A.fs] = __fsub_rn(v[u][1], t1);
                            if (m2) p[2 * A.fs] = __fsub_rn(v[u][2], t2);
                        }
                    }
                }
            }
            __syncthreads();  // sm.tr / sm.mv and the ring are reused by the next item
        }
    }
}

// small groups: one thread per (group, frame), frame fastest
__device__ __forceinline__ void wrap_small_groups(const WrapArgs &A, long long tid) {
    if (tid >= A.n_ranges * A.F) return;
    const long long g = tid / A.F, f = tid - g * A.F;
    const long long s = A.groups[g];
    const int count = (int)max(min((long long)A.groups[g + 1] - s, (long long)(WRAP_SMALL + 1)), -1ll);
    if (count <= 0 || count > WRAP_SMALL) return;  // empty range: nothing to move (pyx:127,141 loop over nothing)
    float *base = A.coords + s * 3 * A.fs + f;
    float v[WRAP_SMALL][3];
#pragma unroll
    for (int k = 0; k < WRAP_SMALL; ++k)
        if (k < count) {
#pragma unroll
            for (int i = 0; i < 3; ++i) v[k][i] = base[(long long)(k * 3 + i) * A.fs];
        }
    float c[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < WRAP_SMALL; ++k)
        if (k < count) {
#pragma unroll
            for (int i = 0; i < 3; ++i) c[i] = __fadd_rn(c[i], __fdiv_rn(__fsub_rn(v[k][i], c[i]), (float)(k + 1)));
        }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float tr;
        if (wrap_decide(A, c[i], i, f, &tr)) {
#pragma unroll
            for (int k = 0; k < WRAP_SMALL; ++k)
                if (k < count) base[(long long)(k * 3 + i) * A.fs] = __fsub_rn(v[k][i], tr);
        }
    }
}

}  // namespace mkb

using namespace mkb;

extern "C" int mkb_wrap_box(mkb_handle_t h, void *stream, const mkb_traj *t, const uint32_t *groups, int64_t n_groups,
                            const uint32_t *centersel, int64_t n_centersel, const float *center) {
    MKB_ENTER(h);
    cudaStream_t st = (cudaStream_t)stream;
    MKB_STREAM_ORDER(h, st);
    if (!t) return fail(h, MKB_ERR_BAD_ARG, "null trajectory");
    if (t->n_atoms < 0 || t->n_frames < 0 || n_groups < 0 || n_centersel < 0)
        return fail(h, MKB_ERR_BAD_ARG, "negative size");
    if (t->n_atoms >= (1ll << 31) || n_centersel >= (1ll << 31) || n_groups >= (1ll << 31))
        return fail(h, MKB_ERR_BAD_ARG, "n_atoms / n_centersel / n_groups must be < 2^31");
    if (n_centersel == 0 && !center) return fail(h, MKB_ERR_BAD_ARG, "center is required when centersel is empty");
    if (n_centersel > 0 && !centersel) return fail(h, MKB_ERR_BAD_ARG, "null centersel");
    const long long F = t->n_frames, n_ranges = n_groups - 1;
    if (F == 0 || n_ranges <= 0 || t->n_atoms == 0) return MKB_OK;  // pyx:110,123: empty loops
    if (!t->coords || !t->box || !groups) return fail(h, MKB_ERR_BAD_ARG, "null argument");
    if (t->frame_stride < F || t->frame_stride_box < F) return fail(h, MKB_ERR_BAD_ARG, "frame stride < n_frames");
    const long long nchunks = cdiv(F, 32);
    const long long small_blocks = cdiv(n_ranges * F, WRAP_THREADS);
    if (small_blocks + 4 * 148 >= (1ll << 31) || nchunks >= (1ll << 31))
        return fail(h, MKB_ERR_BAD_ARG, "groups x frames too large for one launch");
    float *centre = nullptr;
    unsigned *long_list = nullptr, *n_long = nullptr;
    int rc;
    if ((rc = scratch_get(h, S_COM, (size_t)(3 * F), &centre))) return rc;
    if ((rc = scratch_get(h, S_ITEM_CELL, (size_t)n_ranges, &long_list))) return rc;
    if ((rc = scratch_get(h, S_CELL_COUNT, (size_t)1, &n_long))) return rc;
    if (h->timing) MKB_CUDA(h, cudaEventRecord(h->ev[0], st));
    MKB_CUDA(h, cudaMemsetAsync(n_long, 0, sizeof(unsigned), st));
    wrap_classify_kernel<<<(unsigned)cdiv(n_ranges, 256), 256, 0, st>>>(groups, n_ranges, long_list, n_long);
    MKB_LAUNCHED(h);
    if (n_centersel > 0) {
        wrap_center_kernel<<<(unsigned)nchunks, WRAP_THREADS, 0, st>>>(t->coords, F, t->frame_stride, centersel,
                                                                      n_centersel, centre);
        MKB_LAUNCHED(h);
    }
    WrapArgs A;
    A.coords = const_cast<float *>(t->coords);
    A.box = t->box;
    A.F = F; A.fs = t->frame_stride; A.fsb = t->frame_stride_box;
    A.groups = groups; A.n_ranges = n_ranges;
    A.centre = centre;
    A.cx = center ? center[0] : 0.f; A.cy = center ? center[1] : 0.f; A.cz = center ? center[2] : 0.f;
    A.fixed_centre = n_centersel == 0 ? 1 : 0;
    A.long_list = long_list; A.n_long = n_long;
    A.n_long_ctas = (int)std::min<long long>(n_ranges * nchunks, 4ll * h->sm_count);
    if (h->timing) MKB_CUDA(h, cudaEventRecord(h->ev[1], st));
    wrap_groups_kernel<<<(unsigned)(A.n_long_ctas + small_blocks), WRAP_THREADS, 0, st>>>(A);
    MKB_LAUNCHED(h);
    if (h->timing) MKB_CUDA(h, cudaEventRecord(h->ev[2], st));
    return MKB_OK;
}
