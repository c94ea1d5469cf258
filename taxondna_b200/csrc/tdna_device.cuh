// Device-side arithmetic shared by every kernel: plane algebra per distance mode, the fp64
// distance formulas in the reference's exact evaluation order, and StrictMath.log.
//
// Reference lines (src/main/java/com/ggvaidya/TaxonDNA/Common/DNA/Sequence.java):
//   getSharedLength :1414-1485, identical :1250-1303, countTransversions :1352-1371,
//   getK2PDistance :1498-1573, getPairwiseNoBuffer :1699-1717.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tdna {

// ---- symbol classes --------------------------------------------------------------------
// One bit per per-site predicate; a plane set picks the bits a mode needs.
enum SymBit : uint32_t {
    SB_A = 1u << 0,   // IUPAC mask has A   (getint & 1, Sequence.java:1125-1143)
    SB_C = 1u << 1,   // ... C
    SB_G = 1u << 2,   // ... G
    SB_T = 1u << 3,   // ... T
    SB_D = 1u << 4,   // internal gap '-'
    SB_V = 1u << 5,   // letter or '-'  (not '?', not '_')
    SB_U = 1u << 6,   // unambiguous purine/pyrimidine class: A G R C T Y
    SB_PU = 1u << 7,  // purine A G R
    SB_C0 = 1u << 8,  // K2P within-class code bit 0: G or T
    SB_C1 = 1u << 9,  // K2P within-class code bit 1: R or Y
    SB_LET = 1u << 10 // any IUPAC letter
};

// Kernel-facing mode ids (mode x ambiguousBasesAllowed folded together)
enum KMode : int { KM_P_AMBIG = 0, KM_P_NOAMBIG = 1, KM_K2P = 2, KM_TRANS = 3 };

template <int KM> struct ModeTraits;
// accumulators are two 16-bit counts packed in one u32 (L <= 65535)
template <> struct ModeTraits<KM_P_AMBIG> {   // planes: A C G T D V ; acc0 = ident | shared<<16
    static constexpr int P = 6, NACC = 1;
    static constexpr uint32_t bits[6] = {SB_A, SB_C, SB_G, SB_T, SB_D, SB_V};
};
template <> struct ModeTraits<KM_P_NOAMBIG> { // planes: LET D V   ; acc0 = ident | shared<<16
    static constexpr int P = 3, NACC = 1;
    static constexpr uint32_t bits[3] = {SB_LET, SB_D, SB_V};
};
template <> struct ModeTraits<KM_K2P> {       // planes: U PU C0 C1 LET ; acc0 = n | shared<<16, acc1 = ts | tv<<16
    static constexpr int P = 5, NACC = 2;
    static constexpr uint32_t bits[5] = {SB_U, SB_PU, SB_C0, SB_C1, SB_LET};
};
template <> struct ModeTraits<KM_TRANS> {     // planes: U PU V D  ; acc0 = transv | shared<<16
    static constexpr int P = 4, NACC = 1;
    static constexpr uint32_t bits[4] = {SB_U, SB_PU, SB_V, SB_D};
};

// One 32-site word of one pair: x[p], y[p] are the P plane words of the two sequences.
template <int KM>
__device__ __forceinline__ void accumulate_word(const uint32_t *x, const uint32_t *y, uint32_t &a0, uint32_t &a1) {
    if constexpr (KM == KM_P_AMBIG) {
        // ident: IUPAC sets intersect, or both '-'  (identical(), :1262-1302)
        uint32_t id = (x[0] & y[0]) | (x[1] & y[1]) | (x[2] & y[2]) | (x[3] & y[3]) | (x[4] & y[4]);
        uint32_t sh = x[5] & y[5];  // shared: neither '?' nor '_' (:1429-1455,1480)
        a0 += __popc(id) + (__popc(sh) << 16);
    } else if constexpr (KM == KM_P_NOAMBIG) {
        // ambiguousBasesAllowed == false: every letter pair matches (:1280-1283 is always true)
        uint32_t id = (x[0] & y[0]) | (x[1] & y[1]);
        uint32_t sh = x[2] & y[2];
        a0 += __popc(id) + (__popc(sh) << 16);
    } else if constexpr (KM == KM_K2P) {
        uint32_t t = x[0] & y[0];                 // both in {A,G,R,C,T,Y}: n (:1514-1528)
        uint32_t pd = x[1] ^ y[1];                // purine vs pyrimidine
        uint32_t cd = (x[2] ^ y[2]) | (x[3] ^ y[3]);  // different symbol within the class
        uint32_t tv = t & pd;                     // transversion (:1538-1541)
        uint32_t ts = t & ~pd & cd;               // same class, different char (:1535-1537)
        uint32_t sh = x[4] & y[4];                // K2P gate: any two letters (:1457-1471)
        a0 += __popc(t) + (__popc(sh) << 16);
        a1 += __popc(ts) + (__popc(tv) << 16);
    } else {
        uint32_t t = x[0] & y[0];
        uint32_t tv = t & (x[1] ^ y[1]);                     // countTransversions (:1365)
        uint32_t sh = t | (x[2] & y[2] & (x[3] | y[3]));     // (:1446-1451,1472-1478)
        a0 += __popc(tv) + (__popc(sh) << 16);
    }
}

struct Counts {
    int shared, c1, c2, c3;
};

template <int KM> __device__ __forceinline__ Counts unpack_counts(uint32_t a0, uint32_t a1) {
    Counts c;
    c.shared = (int)(a0 >> 16);
    c.c1 = (int)(a0 & 0xffffu);
    if constexpr (KM == KM_K2P) {
        c.c2 = (int)(a1 & 0xffffu);
        c.c3 = (int)(a1 >> 16);
    } else {
        c.c2 = 0;
        c.c3 = 0;
    }
    return c;
}

// ---- StrictMath.log ----------------------------------------------------------------------
// The fdlibm e_log.c algorithm (what java.lang.StrictMath.log is defined as), written with
// explicit round-to-nearest intrinsics so that nvcc can never contract a*b+c into an FMA: the
// result is bit-identical to the same sequence of IEEE operations on the host (oracle/tdna_oracle.c).
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ double strict_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                 Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01,
                 Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    int hx = __double2hiint(x);
    unsigned lx = (unsigned)__double2loint(x);
    int k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -__longlong_as_double(0x7ff0000000000000LL);
        if (hx < 0) return __longlong_as_double(0x7ff8000000000000LL);
        k -= 54;
        x = dmul(x, two54);
        hx = __double2hiint(x);
    }
    if (hx >= 0x7ff00000) return dadd(x, x);
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int i = (hx + 0x95f64) & 0x100000;
    x = __hiloint2double(hx | (i ^ 0x3ff00000), __double2loint(x));
    k += (i >> 20);
    double f = dsub(x, 1.0);
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            double dk = (double)k;
            return dadd(dmul(dk, ln2_hi), dmul(dk, ln2_lo));
        }
        double R = dmul(dmul(f, f), dsub(0.5, dmul(0.33333333333333333, f)));
        if (k == 0) return dsub(f, R);
        double dk = (double)k;
        return dsub(dmul(dk, ln2_hi), dsub(dsub(R, dmul(dk, ln2_lo)), f));
    }
    double s = ddiv(f, dadd(2.0, f));
    double dk = (double)k;
    double z = dmul(s, s);
    i = hx - 0x6147a;
    double w = dmul(z, z);
    int j = 0x6b851 - hx;
    double t1 = dmul(w, dadd(Lg2, dmul(w, dadd(Lg4, dmul(w, Lg6)))));
    double t2 = dmul(z, dadd(Lg1, dmul(w, dadd(Lg3, dmul(w, dadd(Lg5, dmul(w, Lg7)))))));
    i |= j;
    double R = dadd(t2, t1);
    if (i > 0) {
        double hfsq = dmul(dmul(0.5, f), f);
        if (k == 0) return dsub(f, dsub(hfsq, dmul(s, dadd(hfsq, R))));
        return dsub(dmul(dk, ln2_hi), dsub(dsub(hfsq, dadd(dmul(s, dadd(hfsq, R)), dmul(dk, ln2_lo))), f));
    }
    if (k == 0) return dsub(f, dmul(s, dsub(f, R)));
    return dsub(dmul(dk, ln2_hi), dsub(dsub(dmul(s, dsub(f, R)), dmul(dk, ln2_lo)), f));
}

// out-of-line copy for the tile epilogue (64 pairs per thread: inlining 128 logs would bloat the kernel)
__device__ __noinline__ double strict_log_call(double x) { return strict_log(x); }

// ---- distance from counts -----------------------------------------------------------------
// getPairwiseNoBuffer (:1699-1717): gate on the mode-specific shared length, then the mode's
// formula in Java's left-to-right order; -1.0 = inadequate overlap.
template <int KM> __device__ __forceinline__ double distance_from_counts(const Counts &c, int min_overlap) {
    if (c.shared < min_overlap) return -1.0;
    if constexpr (KM == KM_P_AMBIG || KM == KM_P_NOAMBIG) {
        return dsub(1.0, ddiv((double)c.c1, (double)c.shared));  // :1711
    } else if constexpr (KM == KM_TRANS) {
        return ddiv((double)c.c1, (double)c.shared);  // :1708
    } else {
        double n = (double)c.c1, ts = (double)c.c2, tv = (double)c.c3;
        if (c.c2 == 0 && c.c3 == 0 && c.c1 > 0) return 0.0;  // w1 = w2 = 1: log(1) = 0 exactly -> -0 - 0 -> folded to +0 (:1569)
        double w1 = dsub(dsub(1.0, ddiv(dmul(2.0, ts), n)), ddiv(tv, n));  // :1559
        double w2 = dsub(1.0, ddiv(dmul(2.0, tv), n));                     // :1560
        double d = dsub(dmul(-0.5, strict_log_call(w1)), dmul(0.25, strict_log_call(w2)));  // :1564
        if (d <= 0) return 0.0;                                             // :1569-1570
        return d;
    }
}

// ---- packed-plane addressing ----------------------------------------------------------------
// Global layout: [panel of PR rows][plane][word k][row r in panel] as uint32, i.e. within one
// (panel, plane) the words of 64 rows for one k are contiguous (256 B) and consecutive k follow:
// a K-chunk of one (panel, plane) is ONE contiguous block -> one cp.async.bulk per block.
constexpr int PR = 64;  // rows per panel

__host__ __device__ __forceinline__ size_t plane_index(int64_t row, int plane, int k, int P, int W) {
    int64_t panel = row / PR;
    int r = (int)(row % PR);
    return (((size_t)panel * P + plane) * W + k) * PR + r;
}

// Scalar per-pair counts straight from global planes (tdna_pair, row queries, small lookups).
template <int KM>
__device__ __forceinline__ Counts pair_counts_global(const uint32_t *__restrict__ planes, int64_t i, int64_t j, int W) {
    constexpr int P = ModeTraits<KM>::P;
    uint32_t a0 = 0, a1 = 0;
    for (int k = 0; k < W; k++) {
        uint32_t x[P], y[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
            x[p] = planes[plane_index(i, p, k, P, W)];
            y[p] = planes[plane_index(j, p, k, P, W)];
        }
        accumulate_word<KM>(x, y, a0, a1);
    }
    return unpack_counts<KM>(a0, a1);
}

}  // namespace tdna
