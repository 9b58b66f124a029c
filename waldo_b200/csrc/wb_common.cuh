// Common device helpers of the waldo_b200 kernels (sm_100a).
//
// The same sources also compile as plain C++ with -DWB_HOST_EMU: every CTA is then executed by ONE
// host "thread" (blockDim = 1), all kernels being written with block-/grid-stride loops and using
// only the reduction helpers below for cross-thread work.  That build exists for the unit tests
// (tests/emu) to check index arithmetic and formulas without a GPU; it is never loaded by the
// package (waldo_b200/_lib.py refuses a library whose waldo_has_device_code() is 0).
#pragma once
#include <stdint.h>
#include <math.h>
#include <limits.h>

#ifdef WB_HOST_EMU
// ------------------------------------------------------------------ host emulation shims
#include <algorithm>
#include <cstring>
struct wb_dim3 { unsigned x, y, z; wb_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef wb_dim3 dim3;
static thread_local wb_dim3 threadIdx(0, 0, 0), blockIdx, blockDim, gridDim;
#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __syncthreads() ((void)0)
#define __syncwarp() ((void)0)
#define __ldg(p) (*(p))
typedef void* cudaStream_t;
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
// the emulation build is compiled with -ffp-contract=off, so these stay single IEEE operations
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct double2 { double x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#include <atomic>
extern std::atomic<long long> g_wb_launches;
#define WB_LAUNCH(kern, grid, block, smem, stream, ...)                                     \
  do {                                                                                      \
    ++g_wb_launches;                                                                        \
    wb_dim3 g_ = (grid);                                                                    \
    gridDim = g_; blockDim = wb_dim3(1, 1, 1); threadIdx = wb_dim3(0, 0, 0);                \
    for (unsigned bz_ = 0; bz_ < g_.z; ++bz_)                                               \
      for (unsigned by_ = 0; by_ < g_.y; ++by_)                                             \
        for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) { blockIdx = wb_dim3(bx_, by_, bz_); kern(__VA_ARGS__); } \
  } while (0)
#define WB_CHECK_LAUNCH() 0
#define WB_UNROLL
#define WB_UNROLL_N(n)
#define WB_UNROLL_NA
static thread_local float wb_dyn_smem_buf[96 * 1024];
#define WB_DYN_SMEM(name) float* name = wb_dyn_smem_buf
#else
// ------------------------------------------------------------------ device build
#include <cuda_runtime.h>
#include <atomic>
extern std::atomic<long long> g_wb_launches;   // forward on the main thread, backward on autograd's worker thread
#define WB_LAUNCH(kern, grid, block, smem, stream, ...) \
  do { ++g_wb_launches; kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); } while (0)
#define WB_CHECK_LAUNCH() wb_check_launch(__FILE__, __LINE__)
#define WB_UNROLL _Pragma("unroll")
#define WB_PRAGMA_(x) _Pragma(#x)
#define WB_PRAGMA(x) WB_PRAGMA_(x)
#define WB_UNROLL_N(n) WB_PRAGMA_(unroll n)
// loops over the NA layer slots of a template: fully unrolled (register arrays) for the sparse instantiations,
// rolled (local-memory arrays, small code, few registers) for the rare dense one
#define WB_UNROLL_NA _Pragma("unroll (NA <= 8 ? NA : 1)")

#define WB_DYN_SMEM(name) extern __shared__ __align__(16) float name[]
#endif

// trip count of the loops over the layer slots of a template: all NA slots when unrolled, the live count when rolled
#define WB_NEND (NA <= 8 ? NA : ix.n)
#define WB_MAX_L 17
// Which per-warp layer-slot instantiations a kernel carries: 3 = {4 unrolled, 8 unrolled, rolled}, 2 = {4 unrolled, rolled},
// 1 = {rolled}.  Fewer variants = smaller code = fewer instruction-cache misses.  Measured on B200 (profiles/): the forward
// kernels are fastest with all three, the (much larger) backward kernels with the rolled body only.
#ifndef WB_NA_VARIANTS_FWD
#define WB_NA_VARIANTS_FWD 3
#endif
#ifndef WB_NA_VARIANTS_BWD
#define WB_NA_VARIANTS_BWD 1
#endif
// resident CTAs per SM requested through __launch_bounds__ (register budget = 65536 / (256 * N)); tuned on B200
#ifndef WB_OCC_LAYERS_FWD
#define WB_OCC_LAYERS_FWD 4
#endif
#ifndef WB_OCC_PREP_FWD
#define WB_OCC_PREP_FWD 4
#endif
#ifndef WB_OCC_GATHER_FWD
#define WB_OCC_GATHER_FWD 4
#endif
#ifndef WB_OCC_LAYERS_BWD
#define WB_OCC_LAYERS_BWD 2
#endif
#ifndef WB_OCC_PREP_BWD
#define WB_OCC_PREP_BWD 2
#endif
#ifndef WB_OCC_GATHER_BWD
#define WB_OCC_GATHER_BWD 3
#endif
#ifndef WB_LANES_MAX_BWD
#define WB_LANES_MAX_BWD 8   // rows with more live layers take the one-lane-per-pixel form
#endif
#ifndef WB_LANES_PREP_BWD
#define WB_LANES_PREP_BWD 0
#endif
// k_gather_bwd: contexts processed together and channel-loop unrolling (B200 A/B runs, profiles/)
#ifndef WB_GB_ASYNC
#define WB_GB_ASYNC 1   // cp.async operand pipeline in the FAST gather backward (3.15 -> 2.18 ms together with explicit REDG)
#endif
#ifndef WB_GB_MERGE
#define WB_GB_MERGE 1   // neighbouring lanes merge coinciding taps before the global reductions
#endif
#ifndef WB_GB_TG
#define WB_GB_TG 2
#endif
#ifndef WB_GB_UNROLL
#define WB_GB_UNROLL 2
#endif
#define WB_MAX_C 24
#define WB_MAX_NL 21
#define WB_MAX_K 256

#define WB_DEV __device__ __forceinline__

// linear thread id / count inside the CTA and over the grid (1-D launches)
WB_DEV int wb_tid() { return (int)threadIdx.x; }
WB_DEV int wb_nthr() { return (int)blockDim.x; }

// ------------------------------------------------------------------ warp / block reductions
// In the emulation build a "warp" is one lane, so every reduction is the identity.
WB_DEV float wb_warp_sum(float v) {
#ifndef WB_HOST_EMU
  WB_UNROLL for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}
WB_DEV int wb_lane() {
#ifdef WB_HOST_EMU
  return 0;
#else
  return (int)(threadIdx.x & 31);
#endif
}
WB_DEV int wb_warp() {
#ifdef WB_HOST_EMU
  return 0;
#else
  return (int)(threadIdx.x >> 5);
#endif
}
// value of lane `src` (all 32 lanes must call it)
WB_DEV int wb_shfl(int v, int src) {
#ifdef WB_HOST_EMU
  return v;
#else
  return __shfl_sync(0xffffffffu, v, src);
#endif
}
// OR over the warp (all 32 lanes must call it)
WB_DEV unsigned wb_warp_or(unsigned m) {
#ifdef WB_HOST_EMU
  return m;
#else
  return __reduce_or_sync(0xffffffffu, m);
#endif
}
WB_DEV double wb_warp_sum(double v) {
#ifndef WB_HOST_EMU
  WB_UNROLL for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}

// Fire-and-forget float reduction into GLOBAL memory.  `atomicAdd` on a pointer whose address space the compiler cannot
// prove (e.g. one read back from a shared-memory pointer table) compiles to the GENERIC form -- a predicated ATOM plus
// shared-memory and generic CAS loops behind ISSPACEP branches -- instead of one REDG: tell it the space.
WB_DEV void wb_red(float* p, float v) {
#ifdef WB_HOST_EMU
  *p += v;
#else
  // (no "memory" clobber: it would pin every load behind the reduction.  `__builtin_assume(__isGlobal(p)); atomicAdd(p, v);`
  //  also yields REDG, but measured 3-4 % slower in the layer kernels on B200, profiles/r1_v21)
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v));
#endif
}

// ------------------------------------------------------------------ deterministic (fixed-point) accumulation
// Opt-in replacement of wb_red for the scatter targets of decode_bwd (include/waldo_b200.h, det_* fields): the addend is
// rounded to a multiple of the call's fixed-point unit and added as a 64-bit INTEGER to the shadow element of *p.  Integer
// additions commute and associate exactly, so the sum does not depend on the order in which the lanes / CTAs arrive.
// sc[0] = scale (a power of two), sc[3] = overflow flag.
#define WB_DET_BITS 26          // fixed-point unit = 2^-26 of the largest upstream gradient magnitude (as a power of two):
                                // 1.5e-8 of it, i.e. fp32-grade for the image gradients (addends <= 2 max|upstream|), with
                                // 2^37 of it as the range of a sum and 2^36 of it as the range of one addend (intermediate
                                // gradients reach 1e6 .. 1e8 x the upstream ones where the fused score `norm` is small)
// The rare path of wb_red_fixed, NOT inlined (it is reached from ~1000 call sites of the deterministic kernels: inlined it grows
// them by a third and costs 5 % through the instruction cache):
//   * NaN / inf / beyond 2^62 units: raise the "gradients invalid" flag once (a plain store from every thread that gets here
//     would hammer one address); k_det_convert turns a raised flag into NaN gradients, so the caller's NaN guard sees it;
//   * a large addend, 2^52 .. 2^62 units (intermediate gradients reach 1e8 x the upstream ones where the fused score `norm`
//     is ~1e-6: seed 1 of the benchmark inputs, profiles/r2/r2_notes.md): add with the old value returned and check that
//     the 64-bit sum did not wrap.
#ifdef WB_HOST_EMU
static inline void wb_red_fixed_rare(unsigned long long* q, float* sc, float x) {
#else
static __device__ __noinline__ void wb_red_fixed_rare(unsigned long long* q, float* sc, float x) {
#endif
  bool bad = !(fabsf(x) < 4.6e18f);
  if (!bad) {
#ifdef WB_HOST_EMU
    const long long iv = llrintf(x), old = (long long)*q;
    *q += (unsigned long long)iv;
    const long long nw = (long long)*q;
#else
    const long long iv = __float2ll_rn(x);
    const long long old = (long long)atomicAdd(q, (unsigned long long)iv);
    const long long nw = (long long)((unsigned long long)old + (unsigned long long)iv);
#endif
    bad = ((old ^ nw) & (iv ^ nw)) < 0;
  }
  if (bad && reinterpret_cast<volatile float*>(sc)[3] == 0.f) sc[3] = 1.f;
}
WB_DEV void wb_red_fixed(const float* base, int64_t* shadow, float* sc, float* p, float v) {
  const float x = v * __ldg(sc);
  unsigned long long* q = reinterpret_cast<unsigned long long*>(shadow + (p - base));
  if (!(fabsf(x) < 4.5e15f)) { wb_red_fixed_rare(q, sc, x); return; }
  // the usual case, up to 2^52 units (2^11 such addends cannot wrap the sum): fire and forget
#ifdef WB_HOST_EMU
  *q += (unsigned long long)llrintf(x);
#else
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(q), "l"((unsigned long long)__float2ll_rn(x)));
#endif
}

// max |x| over a buffer as the bit pattern of a non-negative float (ordered like unsigned integers; a NaN sorts above inf)
__global__ void k_det_absmax(const float* __restrict__ x, long long n, float* sc) {
  unsigned m = 0u;
  for (long long i = (long long)blockIdx.x * wb_nthr() + wb_tid(); i < n; i += (long long)gridDim.x * wb_nthr()) {
    union { float f; unsigned u; } c;
    c.f = __ldg(x + i);
    const unsigned b = c.u & 0x7fffffffu;
    m = b > m ? b : m;
  }
#ifndef WB_HOST_EMU
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) != 0) return;
#endif
  if (m) atomicMax(reinterpret_cast<unsigned*>(sc) + 2, m);
}
__global__ void k_det_clear(float* sc) { if (blockIdx.x == 0 && wb_tid() == 0) { sc[0] = 1.f; sc[1] = 1.f; sc[2] = 0.f; sc[3] = 0.f; } }
__global__ void k_det_scale(float* sc) {
  if (blockIdx.x != 0 || wb_tid() != 0) return;
  const float gmax = sc[2];   // bits of a non-negative float
  float scale = 1.f, inv = 1.f;
  if (!(gmax < 3.0e38f)) sc[3] = 1.f;   // inf / NaN upstream
  else if (gmax > 0.f) {
    int e;
    frexpf(gmax, &e);                   // gmax <= 2^e
    int se = WB_DET_BITS - e;
    se = se > 100 ? 100 : (se < -100 ? -100 : se);
    scale = ldexpf(1.f, se); inv = ldexpf(1.f, -se);
  }
  sc[0] = scale; sc[1] = inv;
}
// shadow -> fp32 (one rounding), grid-stride.  If any addend so far was NaN / inf / out of the fixed-point range (sc[3]),
// the sums are meaningless: the targets are written as NaN, exactly what the float-reduction path would have propagated,
// so that the caller's NaN guard (reference: synthesizer.py:619) fires instead of wrong gradients being applied.
__global__ void k_det_convert(const int64_t* __restrict__ sh, float* __restrict__ dst, long long n, const float* __restrict__ sc) {
  const double inv = (double)sc[1];
  const bool bad = sc[3] != 0.f;
  const float poison = nanf("");
  for (long long i = (long long)blockIdx.x * wb_nthr() + wb_tid(); i < n; i += (long long)gridDim.x * wb_nthr())
    dst[i] = bad ? poison : (float)((double)sh[i] * inv);
}

#ifndef WB_HOST_EMU
// asynchronous 4-byte copies global -> shared (LDGSTS): operands in flight without holding registers
WB_DEV void wb_cp4(float* smem_dst, const float* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
// same with zero fill: `valid` false copies nothing and writes 0.f (src-size operand 0); gsrc must still be a mapped address
WB_DEV void wb_cp4z(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sa), "l"(gsrc), "r"(n) : "memory");
}
// 16-byte form (LDGSTS.128, L2 only): both addresses 16-byte aligned
WB_DEV void wb_cp16z(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gsrc), "r"(n) : "memory");
}
WB_DEV void wb_cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> WB_DEV void wb_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + transaction barrier (mbarrier, SASS SYNCS): a contiguous run of global memory
// lands in shared memory without passing through the load/store pipe of the SM; completion is counted in bytes on an mbarrier.
WB_DEV unsigned wb_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
WB_DEV void wb_mbar_init(unsigned long long* bar, int arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(wb_smem_addr(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
WB_DEV void wb_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {   // one arrival + the number of bytes the copies will deliver
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(wb_smem_addr(bar)), "r"(bytes) : "memory");
}
WB_DEV void wb_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n.reg .pred P1;\nWB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra WB_DONE;\nbra WB_WAIT;\nWB_DONE:\n}\n"
               ::"r"(wb_smem_addr(bar)), "r"(parity) : "memory");
}
// `bytes` (a multiple of 16) from 16-byte aligned global memory to 16-byte aligned shared memory
WB_DEV void wb_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(wb_smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(wb_smem_addr(bar)) : "memory");
}
WB_DEV void wb_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// L1 prefetch hint (no register, no dependency): used to pull the NEXT context's low-res flow cells while the current
// context is processed
WB_DEV void wb_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#ifndef WB_PF_FLO
#define WB_PF_FLO 0   // measured on B200 (profiles/r1_v44_prefetch_ab.log): SLOWER in both lanes kernels (fwd 0.884 -> 0.912 ms,
#endif                // bwd 1.977 -> 2.066 ms) -- they are bound by issue slots, not by the latency of these L2-resident cells

// position of the nth (0-based) set bit of m  (lanes-per-layer kernels: layer of a slot)
WB_DEV int wb_nth_bit(unsigned m, int nth) { return (int)__fns(m, 0, nth + 1); }
#endif

// ------------------------------------------------------------------ storage types of the HD activations
// `input`, `raw_output` and `out_full` are stored either as fp32 (the reference's precision; parity rules 1-3) or as bf16
// (waldo_decode_fwd_t.storage = 1: half the HBM bytes of the big HD streams, fp32 arithmetic throughout; forward / inference
// only; parity rule 4, tolerance stated in tests/parity.py).  `alpha` -- from which the flow is computed -- stays fp32, so the
// sampling positions do not depend on the storage type.  Kernels are templates over the storage type ST.
struct wb_bf16 { unsigned short x; };
WB_DEV float wb_lds(const float* p) { return __ldg(p); }
WB_DEV float wb_lds(const wb_bf16* p) {
  union { unsigned u; float f; } c;
  c.u = ((unsigned)__ldg(reinterpret_cast<const unsigned short*>(p))) << 16;   // bf16 -> fp32 is exact
  return c.f;
}
WB_DEV void wb_sts(float* p, float v) { *p = v; }
WB_DEV void wb_sts(wb_bf16* p, float v) {   // round to nearest even (NaN stays NaN)
  union { unsigned u; float f; } c;
  c.f = v;
  unsigned r = c.u + 0x7fffu + ((c.u >> 16) & 1u);
  if ((c.u & 0x7fffffffu) > 0x7f800000u) r = c.u | 0x00400000u;
  p->x = (unsigned short)(r >> 16);
}

// 128-bit read-only load of four consecutive floats (16-byte aligned): LDG.E.128
WB_DEV float4 wb_ld4f(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------ ATen-exact bilinear pieces
// grid_sampler_2d (bilinear, zeros, align_corners=False), SURVEY.md Appendix C.  The association
// below -- weights as single products, value accumulated nw -> ne -> sw -> se with fused
// multiply-adds -- reproduces the ATen CPU kernel BIT-EXACTLY (probed: oracle/aten_probe notes in
// DESIGN.md), which is what makes the thresholded / rounded maps of the path index-exact.
struct WbTaps {
  int x0, y0;          // north-west tap (may be out of range)
  float nw, ne, sw, se;
  float ix, iy;        // un-normalised sample position
  float wx0, wx1, wy0, wy1;   // 1-D weights (backward: d/d ix, d/d iy)
};

WB_DEV WbTaps wb_taps(float gx, float gy, int W, int H) {
  WbTaps t;
  // (x * 0.5f is the same correctly rounded value as ATen's x / 2 for every x, without the division sequence)
  t.ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 0.5f);
  t.iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 0.5f);
  float fx = floorf(t.ix), fy = floorf(t.iy);
  float wx1 = __fsub_rn(t.ix, fx), wx0 = __fsub_rn(__fadd_rn(fx, 1.f), t.ix);
  float wy1 = __fsub_rn(t.iy, fy), wy0 = __fsub_rn(__fadd_rn(fy, 1.f), t.iy);
  t.wx0 = wx0; t.wx1 = wx1; t.wy0 = wy0; t.wy1 = wy1;
  t.nw = __fmul_rn(wx0, wy0); t.ne = __fmul_rn(wx1, wy0);
  t.sw = __fmul_rn(wx0, wy1); t.se = __fmul_rn(wx1, wy1);
  // clamp before the int conversion so that wild coordinates (the 2W sentinel is fine, NaN is not
  // expected) cannot overflow; anything outside [-1, size] is out of range for both taps anyway
  fx = fminf(fmaxf(fx, -2.f), (float)W + 1.f);
  fy = fminf(fmaxf(fy, -2.f), (float)H + 1.f);
  t.x0 = (int)fx; t.y0 = (int)fy;
  return t;
}

// 4-bit validity mask of the taps: bit0 nw, bit1 ne, bit2 sw, bit3 se
WB_DEV int wb_tap_mask(const WbTaps& t, int W, int H) {
  int xa = (t.x0 >= 0) & (t.x0 <= W - 1), xb = (t.x0 + 1 >= 0) & (t.x0 + 1 <= W - 1);
  int ya = (t.y0 >= 0) & (t.y0 <= H - 1), yb = (t.y0 + 1 >= 0) & (t.y0 + 1 <= H - 1);
  return (xa & ya) | ((xb & ya) << 1) | ((xa & yb) << 2) | ((xb & yb) << 3);
}

WB_DEV float wb_chain(float vnw, float vne, float vsw, float vse, const WbTaps& t) {
  return __fmaf_rn(vse, t.se, __fmaf_rn(vsw, t.sw, __fmaf_rn(vne, t.ne, __fmul_rn(vnw, t.nw))));
}

// sample one plane (row-major H x W) with zero padding
WB_DEV float wb_sample(const float* __restrict__ plane, const WbTaps& t, int m, int W) {
  const float* p = plane + (long long)t.y0 * W + t.x0;
  float vnw = (m & 1) ? __ldg(p) : 0.f;
  float vne = (m & 2) ? __ldg(p + 1) : 0.f;
  float vsw = (m & 4) ? __ldg(p + W) : 0.f;
  float vse = (m & 8) ? __ldg(p + W + 1) : 0.f;
  return wb_chain(vnw, vne, vsw, vse, t);
}

// Branch-free form of the four taps.  Two row offsets (rows clamped into the plane, column xL = clamp(x0, 0, W-2)), the
// loads are p0[0], p0[1], p1[0], p1[1]; each tap's validity and weight are folded into four POSITIONAL weights
// (an out-of-range tap gets weight 0 instead of value 0).  Accumulated in position order this is the ATen chain
// nw -> ne -> sw -> se with exact zeros inserted, i.e. the same rounding sequence.  Needs W >= 2.
struct WbTap2 {
  unsigned o0, o1;   // element offsets of (row0, xL) and (row1, xL) inside one H x W plane
  float w[4];        // positional weights: (row0,L) (row0,R) (row1,L) (row1,R)
  int sel;           // where the tap columns sit: 0 x0 == xL (interior) / 1 x0 == -1 / 2 x0 == W-1 / 3 none in range
  int vy;            // bit 0: row y0 in range, bit 1: row y0+1 in range
};
// move four per-tap coefficients (nw, ne, sw, se) to positions, zeroing out-of-range taps
WB_DEV void wb_pos4(const WbTap2& a, float c_nw, float c_ne, float c_sw, float c_se, float* o) {
  const float r0 = (a.vy & 1) ? 1.f : 0.f, r1 = (a.vy & 2) ? 1.f : 0.f;
  float l0, rr0, l1, rr1;
  if (a.sel == 0) { l0 = c_nw; rr0 = c_ne; l1 = c_sw; rr1 = c_se; }
  else if (a.sel == 1) { l0 = c_ne; rr0 = 0.f; l1 = c_se; rr1 = 0.f; }
  else if (a.sel == 2) { l0 = 0.f; rr0 = c_nw; l1 = 0.f; rr1 = c_sw; }
  else { l0 = rr0 = l1 = rr1 = 0.f; }
  o[0] = l0 * r0; o[1] = rr0 * r0; o[2] = l1 * r1; o[3] = rr1 * r1;
}
WB_DEV WbTap2 wb_tap2(const WbTaps& t, int W, int H) {
  WbTap2 a;
  const int x0 = t.x0;
  const int xL = min(max(x0, 0), W - 2);
  a.sel = (x0 == xL) ? 0 : (x0 == -1 ? 1 : (x0 == W - 1 ? 2 : 3));
  a.vy = ((t.y0 >= 0 && t.y0 <= H - 1) ? 1 : 0) | ((t.y0 + 1 >= 0 && t.y0 + 1 <= H - 1) ? 2 : 0);
  const int y0 = min(max(t.y0, 0), H - 1), y1 = min(max(t.y0 + 1, 0), H - 1);
  a.o0 = (unsigned)(y0 * W + xL); a.o1 = (unsigned)(y1 * W + xL);
  wb_pos4(a, t.nw, t.ne, t.sw, t.se, a.w);
  return a;
}
// bilinear value through precomputed taps; p0 / p1 point at (row0, xL) / (row1, xL) of the plane
WB_DEV float wb_gather2(const float* __restrict__ p0, const float* __restrict__ p1, const float* w) {
  return __fmaf_rn(__ldg(p1 + 1), w[3], __fmaf_rn(__ldg(p1), w[2], __fmaf_rn(__ldg(p0 + 1), w[1], __fmul_rn(__ldg(p0), w[0]))));
}
// same for a plane that stores 2A-1 while the sampled quantity is A (zero padding applies to A)
template <typename ST>
WB_DEV float wb_gather2_01(const ST* __restrict__ p0, const ST* __restrict__ p1, const float* w) {
  const float v0 = (wb_lds(p0) + 1.f) * 0.5f, v1 = (wb_lds(p0 + 1) + 1.f) * 0.5f;
  const float v2 = (wb_lds(p1) + 1.f) * 0.5f, v3 = (wb_lds(p1 + 1) + 1.f) * 0.5f;
  return __fmaf_rn(v3, w[3], __fmaf_rn(v2, w[2], __fmaf_rn(v1, w[1], __fmul_rn(v0, w[0]))));
}

// upsample_bilinear2d (align_corners=False) source coordinates along one axis, ATen association:
// src = max(r*(dst+0.5)-0.5, 0); i0 = int(src); i1 = min(i0+1, n-1); l1 = src-i0; l0 = 1-l1.
struct WbAxis { int i0, i1; float l0, l1; };
WB_DEV WbAxis wb_axis(int dst, float r, int n_in) {
  WbAxis a;
  float s = fmaxf(__fsub_rn(__fmul_rn(r, __fadd_rn((float)dst, 0.5f)), 0.5f), 0.f);
  a.i0 = min((int)s, n_in - 1);
  a.i1 = min(a.i0 + 1, n_in - 1);
  a.l1 = fminf(fmaxf(__fsub_rn(s, (float)a.i0), 0.f), 1.f);
  a.l0 = __fsub_rn(1.f, a.l1);
  return a;
}
// value = fma(row0, ly0, row1*ly1), row = fma(v[x0], lx0, v[x1]*lx1)  (bit-exact with ATen CPU)
WB_DEV float wb_lerp2(float v00, float v01, float v10, float v11, const WbAxis& ax, const WbAxis& ay) {
  float r0 = __fmaf_rn(v00, ax.l0, __fmul_rn(v01, ax.l1));
  float r1 = __fmaf_rn(v10, ax.l0, __fmul_rn(v11, ax.l1));
  return __fmaf_rn(r0, ay.l0, __fmul_rn(r1, ay.l1));
}
