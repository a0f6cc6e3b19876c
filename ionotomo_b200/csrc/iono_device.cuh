// Device-side building blocks shared by the ray-integral kernels (sm_100a).
//
//  * Axis tables and cell lookup  == SciPy RegularGridInterpolator(method='linear')
//    index/weight semantics (reference: geometry/tri_cubic.py:22,69-75;
//    SURVEY.md Appendix A.1).
//  * Simpson weights              == scipy.integrate.simps(y, x, even='avg')
//    (reference: inversion/forward_equation.py:28; restated by the reference in
//    tomography/integrate.py:50-153), as a per-sample
//    weight so that forward (gather) and adjoint (scatter) share one rule.
//  * Warp-private TMA bulk-copy ring for streaming the materialised rays.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace iono {

// ---------------------------------------------------------------------------
// Grid description passed by value to kernels.
// tab[i] = { g[i], 1/(g[i+1]-g[i]) } for i < n-1, tab[n-1] = { g[n-1], 0 }.
// ---------------------------------------------------------------------------
struct Axis {
    const double2 *tab;  // device
    int n;
    int uniform;      // 1: direct cell index, 0: bisection
    int exact;        // 1: nodes equal g0 + i*d to a few ulps (np.linspace): in-cell coordinate by arithmetic
    double inv_d;     // (n-1)/(g[n-1]-g[0])               (uniform only)
    double c_guess;   // -g[0]*inv_d - 0.5                 (uniform only)
    double g0, glast;
};

struct Grid {
    Axis ax[3];
};

// 2^52 + 2^51: adding it rounds a double in (-2^31, 2^31) to the nearest integer,
// which then sits in the low 32 bits of the mantissa.
#define IONO_MAGIC 6755399441055744.0

// Cell index i and in-cell coordinate t of SciPy's RGI for one axis:
//   i = clip(searchsorted(g, x, 'right') - 1, 0, n-2),  t = (x - g[i]) / (g[i+1] - g[i]).
// `tab` may point to shared or global memory.  oob |= (x < g[0] || x > g[n-1] || isnan(x)).
template <bool UNIFORM>
__device__ __forceinline__ void locate(const double2 *__restrict__ tab, const Axis &a, double x,
                                       int &i, double &t, bool &oob) {
    if (UNIFORM) {
        double v = fma(x, a.inv_d, a.c_guess);               // (x-g0)/d - 0.5
        i = __double2loint(v + IONO_MAGIC);                   // ~floor((x-g0)/d)
        i = min(max(i, 0), a.n - 2);
    } else {
        int lo = 0, hi = a.n - 1;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (x >= tab[mid].x) lo = mid; else hi = mid;
        }
        i = min(lo, a.n - 2);
    }
    double2 e = tab[i];
    t = (x - e.x) * e.y;
    if (!(t >= 0.0 && t < 1.0)) {
        // Rare: guess off by one, x on/over the last node, outside the grid, or NaN.
        while (i > 0 && x < tab[i].x) --i;
        while (i < a.n - 2 && x >= tab[i + 1].x) ++i;
        e = tab[i];
        t = (x - e.x) * e.y;
        oob = oob || !(x >= a.g0 && x <= a.glast);
    }
}

// 1/d to ~1 ulp: hardware seed (rcp.approx.ftz.f64, ~2^-23) + two Newton steps.
// d is a product of Simpson interval lengths: finite, normal, non-zero for any
// strictly monotone s; d == 0 yields inf/NaN like the reference's division.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// Simpson 'avg' weight of sample i of an N-sample ray, given the abscissae
// s[i-2..i+2] (entries outside [0,N) are never used).
//   odd N : composite Simpson, triples start at 0,2,..,N-3 (non-uniform 3-point form)
//   even N: 1/2 {Simpson on points 0..N-2 + trapezoid on the last interval}
//         + 1/2 {trapezoid on the first interval + Simpson on points 1..N-1}
// With hm=h[i-1], hp=h[i], h0=h[i-2], h3=h[i+1] every term shares 1/(6 hm hp):
//   middle of a triple : (hm+hp)^3
//   right end          : (h0+hm)(2hm-h0) hp
//   left end           : (hp+h3)(2hp-h3) hm
// The part that depends on (i, N) only is split off so that a loop over rays with i fixed evaluates it once
// (iono_adjoint_runs.cuh); simpson_weight() is the two halves back to back, same operations in the same order.
struct SimpsonCoef {
    double cm, cr, cl, tp, tm;        // triple-middle / right-end / left-end factors, trapezoid multipliers of hp, hm
    bool first, last, has_h0, has_h3;
};
__device__ __forceinline__ SimpsonCoef simpson_coef(int i, int N, bool n_odd) {
    SimpsonCoef k;
    k.first = i < 1; k.last = i > N - 2; k.has_h0 = i >= 2; k.has_h3 = i <= N - 3;
    const int par = i & 1;
    k.tp = 0.0; k.tm = 0.0;
    if (n_odd) {   // warp-uniform branch
        k.cm = par ? 1.0 : 0.0;
        k.cr = (par | (i < 2)) ? 0.0 : 1.0;
        k.cl = (par | (i > N - 3)) ? 0.0 : 1.0;
    } else {
        k.cm = (par ? (i <= N - 3) : (i >= 2)) ? 0.5 : 0.0;
        k.cr = (i >= 2 + par) ? 0.5 : 0.0;
        k.cl = (i <= N - 4 + par) ? 0.5 : 0.0;
        if (i < 2 || i > N - 3) {
            k.tp = (double)((i == 0) + (i == N - 2));
            k.tm = (double)((i == 1) + (i == N - 1));
        }
    }
    return k;
}
__device__ __forceinline__ double simpson_weight_c(const SimpsonCoef &k, double sm2, double sm1, double s0, double sp1,
                                                   double sp2) {
    double hm = s0 - sm1, hp = sp1 - s0;
    if (k.first) hm = hp;
    if (k.last) hp = hm;
    const double h0 = k.has_h0 ? sm1 - sm2 : hm;
    const double h3 = k.has_h3 ? sp2 - sp1 : hp;
    const double trap = 0.25 * (hp * k.tp + hm * k.tm);      // 0 away from the ends and for odd N
    const double A = hm + hp;
    double num = (k.cm * A) * (A * A);
    num = fma(k.cr * (h0 + hm) * fma(2.0, hm, -h0), hp, num);
    num = fma(k.cl * (hp + h3) * fma(2.0, hp, -h3), hm, num);
    const double r = fast_rcp(6.0 * hm * hp);
    return fma(num, r, trap);
}
__device__ __forceinline__ double simpson_weight(int i, int N, bool n_odd, double sm2, double sm1, double s0,
                                                 double sp1, double sp2) {
    return simpson_weight_c(simpson_coef(i, N, n_odd), sm2, sm1, s0, sp1, sp2);
}

// ---------------------------------------------------------------------------
// mbarrier + TMA 1-D bulk copy (cp.async.bulk): global -> shared, completion by
// transaction bytes on an mbarrier.  Used as a warp-private ring: the warp's
// lane 0 is the producer, all 32 lanes are consumers, so no "empty" barrier is
// needed (a __syncwarp() orders the last read before the re-fill).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// One lane of the (converged) warp.  With elect.sync the compiler knows that exactly one thread runs the guarded
// block: a bulk copy inside it is a single UBLKCP, not a loop over the active lanes' addresses.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// The same three operations on 32-bit shared-window addresses computed ONCE (smem_u32 of the ring / barrier array at
// kernel start): inside a hot loop the generic->shared conversion of a pointer with a run-time offset is otherwise
// re-derived at every use (S2UR CgaCtaId, ULEA, ... -- a dozen uniform instructions per copy).
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_s(uint32_t dst, const void *src_gmem, uint32_t bytes, uint32_t bar,
                                           uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src_gmem), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}

// streaming read (ray data without the bulk path): do not allocate in L1, evict first from L2
__device__ __forceinline__ double ld_stream(const double *p, uint64_t policy) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;"
                 : "=d"(v)
                 : "l"(p), "l"(policy));
    return v;
}

// ---------------------------------------------------------------------------
// Quad layout of a grid field (forward gathers): record v = (ix*ny + iy)*nz + iz holds
//   { f[ix,iy,iz], f[ix,iy,iz+1], f[ix,iy+1,iz], f[ix,iy+1,iz+1] }        (32 bytes, one sector)
// so the 8 corners of a cell are TWO 256-bit loads (records v and v + ny*nz) instead of eight 64-bit
// ones: 2 instead of 8 L1 requests per sample and a quarter of the tag lookups.  Indices past the
// last node are clamped (those values only ever meet a zero weight).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ld_quad(const double4 *p, double &a, double &b, double &c, double &d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// Ray traversal order.  Work index q in [0, n0*n1*n2) -> ray index
// i0*st0 + i1*st1 + i2*st2 with q = i0 + n0*(i1 + n1*i2): consecutive q (the
// warps of one CTA) walk axis 0 first.
// ---------------------------------------------------------------------------
struct RayOrder {
    int n0, n1, n2;
    int st0, st1, st2;
};
__device__ __forceinline__ long long ray_of(const RayOrder &o, int q) {
    const int i0 = q % o.n0;
    const int r = q / o.n0;
    const int i1 = r % o.n1;
    const int i2 = r / o.n1;
    return (long long)(i0 * o.st0 + i1 * o.st1 + i2 * o.st2);
}

}  // namespace iono
