// common.cuh — shared device/host definitions of the B200 PHMM path.
//
// Numeric contract.  The reference computes in log space (prob.rs:13,181-253: Prob = ln p, `+` is
// logaddexp).  Here every probability is a LINEAR f64 mantissa with a separate power-of-two exponent
// ("XF" = value * 2^e).  Multiplying by 2^k is exact, so results are the IEEE f64 results of the linear
// recurrence with an unbounded exponent range, independent of tiling or launch geometry; the only place a
// term is dropped is when two addends differ by more than 2^1022, where log-space f64 drops it as well
// (x + ln_1p(exp(y-x)) == x for y-x < -745).  Values cross the C ABI as natural logs (ln v + e ln 2).
// Storage: a DP cell keeps m, i, d mantissas and ONE shared int32 exponent (28 B/cell).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define MAX_ACTIVE 400          // hmmv2/table.rs:22 MAX_ACTIVE_NODES
#define N_DEL_ROUNDS 5          // 1 + n_max_gaps (forward.rs:439-465, table.rs:17 MAX_DEL = 4)
#define HALO_HOPS 6             // 1 (M/I step) + 5 (Del rounds) dependency hops per row
#define XF_ZERO_E (-(1 << 30))  // exponent carried by an exact zero

struct XF {
    double v;
    int e;
};

// linear-space parameters (exp of PHMMParams' logs, params.rs:16-66)
struct LinParams {
    double p_mismatch, p_match, p_random, p_end;
    double p_MM, p_IM, p_DM, p_MI, p_II, p_DI, p_MD, p_ID, p_DD;
    uint32_t n_active_nodes, n_warmup, warmup_threshold, n_max_gaps;
    double active_node_max_ratio;
};

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

HD double pow2i(int k) {  // 2^k ; flushes below the normal range, saturates above
    if (k < -1022) return 0.0;
    if (k > 1023) k = 1023;
    union { long long i; double d; } u;
    u.i = (long long)(k + 1023) << 52;
    return u.d;
}
HD int ilogb_pos(double v) {  // floor(log2 v) for v > 0 (denormals report -1023)
    union { long long i; double d; } u;
    u.d = v;
    return (int)((u.i >> 52) & 0x7ff) - 1023;
}
HD XF xf(double v, int e) { XF r; r.v = v; r.e = (v == 0.0) ? XF_ZERO_E : e; return r; }
HD XF xf_zero() { XF r; r.v = 0.0; r.e = XF_ZERO_E; return r; }
HD XF xadd(XF a, XF b) {
    if (a.v == 0.0) return b;
    if (b.v == 0.0) return a;
    XF r;
    if (a.e >= b.e) { r.e = a.e; r.v = a.v + b.v * pow2i(b.e - a.e); }
    else { r.e = b.e; r.v = b.v + a.v * pow2i(a.e - b.e); }
    return r;
}
HD XF xmul(XF a, double c) { return xf(a.v * c, a.e); }
// bring the mantissa into [1,2) (keeps value); zero stays zero
HD XF xnorm(XF a) {
    if (a.v == 0.0) return xf_zero();
    int q = ilogb_pos(a.v);
    if (q <= -1023) {  // denormal mantissa: rescale in two steps
        a.v *= 4503599627370496.0;  // 2^52
        a.e -= 52;
        q = ilogb_pos(a.v);
    }
    XF r; r.v = a.v * pow2i(-q); r.e = a.e + q;
    return r;
}
// total binary exponent floor(log2(value)) ; zero -> XF_ZERO_E
HD int xexp(XF a) { return a.v == 0.0 ? XF_ZERO_E : a.e + ilogb_pos(a.v); }
// natural log of the value (ABI export); zero -> -inf
HD double xlog(XF a) {
    if (a.v == 0.0) return -INFINITY;
    return log(a.v) + (double)a.e * 0.693147180559945309417232121458;
}
// a > b as real numbers (both normalised or not)
HD bool xgt(XF a, XF b) {
    if (b.v == 0.0) return a.v != 0.0;
    if (a.v == 0.0) return false;
    XF x = xnorm(a), y = xnorm(b);
    return x.e > y.e || (x.e == y.e && x.v > y.v);
}

// One DP cell brought to a shared exponent: mantissas <= 2, exponent = that of the largest state.
struct Cell { double m, i, d; int e; };
HD Cell cell_pack(XF m, XF i, XF d) {
    int em = xexp(m), ei = xexp(i), ed = xexp(d);
    int E = em > ei ? em : ei;
    E = E > ed ? E : ed;
    Cell c;
    if (E == XF_ZERO_E) { c.m = c.i = c.d = 0.0; c.e = 0; return c; }
    c.e = E;
    c.m = m.v == 0.0 ? 0.0 : m.v * pow2i(m.e - E);
    c.i = i.v == 0.0 ? 0.0 : i.v * pow2i(i.e - E);
    c.d = d.v == 0.0 ? 0.0 : d.v * pow2i(d.e - E);
    return c;
}

// Row kinds in a row store
enum { ROW_NONE = 0, ROW_DENSE = 1, ROW_SPARSE = 2 };

// Descriptor of one stored DP row (PHMMTable, table.rs:42-73)
struct RowDesc {
    int kind;        // ROW_*
    uint32_t n_ent;  // sparse: merged entries (mi entries first, then d-only entries)
    uint32_t n_mi;   // sparse: entries holding m/i (== |nodes| of the step)
    uint32_t n_d;    // sparse: entries holding d
    uint64_t off;    // dense: slab index ; sparse: byte offset into the arena
    XF mb, ib, e;
};

// sparse row payload layout in the arena (8-byte aligned), for n_ent entries, n_d d-entries:
//   double m[n_ent]; double i[n_ent]; double d[n_ent]; uint32 id[n_ent]; int32 ex[n_ent]; uint16 dlist[n_d]
HD uint64_t sparse_row_bytes(uint32_t n_ent, uint32_t n_d) {
    uint64_t b = (uint64_t)n_ent * 32 + (uint64_t)n_d * 2;
    return (b + 7) & ~(uint64_t)7;
}

// dense slab layout: double m[N], i[N], d[N]; int32 ex[N]   (N padded to a multiple of 2)
struct DenseRowPtr {
    const double *m, *i, *d;
    const int* ex;
};
