// lq_common.cuh -- geometry, HBM layout and 3x3 complex f64 algebra shared by every kernel.
//
// HBM layout ("chunked SoA", f64 complex = double2, every access a 128-bit load/store):
//   a field with NPL planes per site is stored as  F[((p >> 5) * NPL + plane) * 32 + (p & 31)]
//   i.e. chunks of 32 consecutive site slots, and inside a chunk plane-major: a warp reading one plane of 32
//   consecutive slots reads 512 contiguous bytes, and the planes of one site sit at the compile-time stride of
//   512 bytes (immediate offsets in the load instructions; one address computation per 3x3 matrix).
//     links : NPL = 9*D, plane = dir*9 + k,  k = 3*row + col
//     efield: NPL = 4*D, plane = dir*4 + q,  element q = (e_{2q}, e_{2q+1})
//     gauss : NPL = 9
//   slot p of a site: rows of x0 keep the reference's order (x0 fastest, lattice.rs:909-916, over the extents +
//   ghost layers of decomposed directions) but inside a row the even-x0 sites come first, then the odd ones:
//     p = (s - x0) + (x0 & 1) * ceil(ext0/2) + (x0 >> 1),   s = storage-lexicographic site index.
//   The sites of one checkerboard colour inside a row are therefore contiguous: whole-lattice kernels and
//   even/odd sweeps both read/write fully used sectors.
//
// Build modes:  default = CUDA (sm_100a).  -DLQ_HOST_EMU = the same kernel bodies run in plain host loops;
// that build is TEST INFRASTRUCTURE (CPU CI of host logic / world_size-2 gloo tests) and is never loaded by
// the package (see tests/emu.py).  The product library has no CPU path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <float.h>

#ifdef LQ_HOST_EMU
#include <cstdlib>
#include <cstring>
struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
#define LQ_HD inline
#define LQ_NOINLINE static
#define LQ_LDG(p) (*(p))
#define LQ_RESTRICT
#else
#include <cuda_runtime.h>
#define LQ_HD __host__ __device__ __forceinline__
#define LQ_NOINLINE static __host__ __device__ __noinline__
#ifdef __CUDA_ARCH__
#define LQ_LDG(p) __ldg(p)
#else
#define LQ_LDG(p) (*(p))
#endif
#define LQ_RESTRICT __restrict__
#endif

#define LQ_MAXD 4
#define LQ_EPS 2.220446049250313e-16 /* f64::EPSILON */
#define LQ_PI 3.14159265358979323846264338327950288

typedef long long lq_i64;

// ---------------------------------------------------------------------------------------------- geometry
struct LqGeom {
  int D;
  int ext[LQ_MAXD];    // rank-local interior extents
  int sext[LQ_MAXD];   // storage extents = ext + 2*ghost
  int ghost[LQ_MAXD];  // 1: direction is decomposed and carries one-site-deep ghost layers
  int gext[LQ_MAXD];   // global extents
  int goff[LQ_MAXD];   // global coordinate of the local interior origin
  lq_i64 sstride[LQ_MAXD];
  lq_i64 nstride[LQ_MAXD];  // stride used by the neighbour shifts (== sstride; tools/kbench zeroes it for a locality bound)
  lq_i64 gstride[LQ_MAXD];
  lq_i64 lstride[LQ_MAXD];  // strides of the rank-local interior block in reference order (AoS boundary)
  lq_i64 vol;               // interior sites
  lq_i64 svol;              // storage sites
  lq_i64 nchunk;            // 32-slot chunks per field = ceil(svol / 32)
  int ne0;                  // even-x0 sites per row = (ext0+1)/2
  // thread -> site mapping of the tuned kernels: the lattice is walked tile by tile (tile[d] divides ext[d],
  // tile[0] even) so that the sites a thread block touches form a compact 4-D brick (L1 reuse of neighbours).
  int tile[LQ_MAXD];
  int ntile[LQ_MAXD];
  int tvol;
};

template <int D>
struct Site {
  lq_i64 s;  // storage-lexicographic index
  int x[D];  // storage coordinates (interior coordinate + ghost)
};

// slot of a site (row-local parity split) and element address of (slot, plane) in a field of npl planes
template <int D>
LQ_HD lq_i64 lq_slot(const LqGeom& g, const Site<D>& st) {
  const int x0 = st.x[0];
  return (st.s - x0) + (x0 & 1) * g.ne0 + (x0 >> 1);
}
LQ_HD lq_i64 lq_slot_s(const LqGeom& g, lq_i64 s) {  // from the storage index alone (direction 0 is never ghosted)
  const int x0 = (int)(s % g.sext[0]);
  return (s - x0) + (x0 & 1) * g.ne0 + (x0 >> 1);
}
LQ_HD lq_i64 lq_addr(lq_i64 p, int npl, int plane) { return ((p >> 5) * npl + plane) * 32 + (p & 31); }

// n in [0, vol) -> site.  Rows of x0 are walked "even x0 first, then odd x0" so that consecutive n (lanes of
// a warp) touch consecutive memory in both halves of the plane.
// (Ordinals below 2^31 -- every lattice that fits a GPU today -- take 32-bit divisions: a 64-bit division costs the
// GPU some eighty instructions, and a thread of the streaming kernels has only a few hundred of useful work.)
template <int D>
LQ_HD Site<D> lq_site(const LqGeom& g, lq_i64 n) {
  Site<D> st;
  if (n < ((lq_i64)1 << 31)) {
    unsigned row = (unsigned)n / (unsigned)g.ext[0];
    int lane = (int)((unsigned)n - row * (unsigned)g.ext[0]);
    int x0 = lane < g.ne0 ? 2 * lane : 2 * (lane - g.ne0) + 1;
    st.x[0] = x0 + g.ghost[0];
    st.s = st.x[0];
#pragma unroll
    for (int d = 1; d < D; ++d) {
      unsigned q = row / (unsigned)g.ext[d];
      int xd = (int)(row - q * (unsigned)g.ext[d]);
      row = q;
      st.x[d] = xd + g.ghost[d];
      st.s += (lq_i64)st.x[d] * g.sstride[d];
    }
    return st;
  }
  lq_i64 row = n / g.ext[0];
  int lane = (int)(n - row * g.ext[0]);
  int x0 = lane < g.ne0 ? 2 * lane : 2 * (lane - g.ne0) + 1;
  st.x[0] = x0 + g.ghost[0];
  st.s = st.x[0];
#pragma unroll
  for (int d = 1; d < D; ++d) {
    lq_i64 q = row / g.ext[d];
    int xd = (int)(row - q * g.ext[d]);
    row = q;
    st.x[d] = xd + g.ghost[d];
    st.s += (lq_i64)st.x[d] * g.sstride[d];
  }
  return st;
}
// n in [0, vol) -> site, walking the lattice tile by tile; inside a tile x0 runs "even first, then odd".
template <int D>
LQ_HD Site<D> lq_site_tiled(const LqGeom& g, lq_i64 n) {
  Site<D> st;
  lq_i64 tid = n / g.tvol;
  int w = (int)(n - tid * g.tvol);
  int h0 = g.tile[0] >> 1;
  int w0 = w % g.tile[0];
  w /= g.tile[0];
  int x0l = w0 < h0 ? 2 * w0 : 2 * (w0 - h0) + 1;
  int t0 = (int)(tid % g.ntile[0]);
  tid /= g.ntile[0];
  st.x[0] = t0 * g.tile[0] + x0l + g.ghost[0];
  st.s = st.x[0];
#pragma unroll
  for (int d = 1; d < D; ++d) {
    int wd = w % g.tile[d];
    w /= g.tile[d];
    int td = (int)(tid % g.ntile[d]);
    tid /= g.ntile[d];
    st.x[d] = td * g.tile[d] + wd + g.ghost[d];
    st.s += (lq_i64)st.x[d] * g.sstride[d];
  }
  return st;
}
// checkerboard enumeration: n in [0, vol/2) -> the n-th site of colour `parity` (global coordinate sum & 1).
// Needs every extent even.
template <int D>
LQ_HD Site<D> lq_site_eo(const LqGeom& g, lq_i64 n, int parity) {
  Site<D> st;
  int h0 = g.ext[0] >> 1;
  int psum = parity;
  st.s = 0;
  if (n < ((lq_i64)1 << 31)) {  // 32-bit divisions, see lq_site
    unsigned row = (unsigned)n / (unsigned)h0;
    int k = (int)((unsigned)n - row * (unsigned)h0);
#pragma unroll
    for (int d = 1; d < D; ++d) {
      unsigned q = row / (unsigned)g.ext[d];
      int xd = (int)(row - q * (unsigned)g.ext[d]);
      row = q;
      psum += xd + g.goff[d];
      st.x[d] = xd + g.ghost[d];
      st.s += (lq_i64)st.x[d] * g.sstride[d];
    }
    int x0 = 2 * k + ((psum + g.goff[0]) & 1);
    st.x[0] = x0 + g.ghost[0];
    st.s += st.x[0];
    return st;
  }
  lq_i64 row = n / h0;
  int k = (int)(n - row * h0);
#pragma unroll
  for (int d = 1; d < D; ++d) {
    lq_i64 q = row / g.ext[d];
    int xd = (int)(row - q * g.ext[d]);
    row = q;
    psum += xd + g.goff[d];
    st.x[d] = xd + g.ghost[d];
    st.s += (lq_i64)st.x[d] * g.sstride[d];
  }
  int x0 = 2 * k + ((psum + g.goff[0]) & 1);
  st.x[0] = x0 + g.ghost[0];
  st.s += st.x[0];
  return st;
}
// Site of one sweep sub-step.  Even extents (odd_mask == 0): `n` enumerates the vol/2 sites of colour `parity`.
// Odd extents (non-decomposed contexts; the reference's sequential sweeps accept any N >= 2): two colours are not
// enough on a periodic ring of odd length, so the colour of a site is (boundary mask, parity) with bit d of the mask set
// iff ext[d] is odd and x_d = ext[d] - 1.  Neighbours x, x + nu always differ: in parity away from the wrap (and for
// even ext[nu]), in bit nu of the mask at x_nu = N-2 -> N-1 and N-1 -> 0.  Here `n` enumerates ALL sites and the call
// returns false for the ones outside the class (odd lattices are small: the scan is cheaper than a second enumeration).
template <int D>
LQ_HD bool lq_site_class(const LqGeom& g, lq_i64 n, int parity, int odd_mask, int cmask, Site<D>& st) {
  if (odd_mask == 0) {
    st = lq_site_eo<D>(g, n, parity);
    return true;
  }
  if (n >= g.vol) return false;
  st = lq_site<D>(g, n);
  int m = 0, ps = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const int xd = st.x[d] - g.ghost[d];
    ps += xd;
    if (((odd_mask >> d) & 1) && xd == g.ext[d] - 1) m |= 1 << d;
  }
  return m == cmask && (ps & 1) == parity;
}
// add_point_direction (lattice.rs:303-323): one step with periodic wrap.  In a ghosted direction interior
// sites never wrap (the neighbour is the ghost layer).
template <int D>
LQ_HD Site<D> lq_up(const LqGeom& g, Site<D> st, int d) {
  if (st.x[d] + 1 < g.sext[d]) {
    st.x[d] += 1;
    st.s += g.nstride[d];
  } else {
    st.s -= (lq_i64)st.x[d] * g.nstride[d];
    st.x[d] = 0;
  }
  return st;
}
template <int D>
LQ_HD Site<D> lq_dn(const LqGeom& g, Site<D> st, int d) {
  if (st.x[d] > 0) {
    st.x[d] -= 1;
    st.s -= g.nstride[d];
  } else {
    st.x[d] = g.sext[d] - 1;
    st.s += (lq_i64)st.x[d] * g.nstride[d];
  }
  return st;
}
// reference-order index of an interior site inside the rank-local block (AoS boundary)
template <int D>
LQ_HD lq_i64 lq_local_index(const LqGeom& g, const Site<D>& st) {
  lq_i64 l = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) l += (lq_i64)(st.x[d] - g.ghost[d]) * g.lstride[d];
  return l;
}
// global reference-order site index (RNG stream id: results do not depend on the decomposition)
template <int D>
LQ_HD lq_i64 lq_global_index(const LqGeom& g, const Site<D>& st) {
  lq_i64 l = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) l += (lq_i64)(st.x[d] - g.ghost[d] + g.goff[d]) * g.gstride[d];
  return l;
}

// ---------------------------------------------------------------------------------------------- complex
typedef double2 cx;
LQ_HD cx cmk(double re, double im) { return make_double2(re, im); }
LQ_HD cx cadd(cx a, cx b) { return cmk(a.x + b.x, a.y + b.y); }
LQ_HD cx csub(cx a, cx b) { return cmk(a.x - b.x, a.y - b.y); }
LQ_HD cx cneg(cx a) { return cmk(-a.x, -a.y); }
LQ_HD cx cconj(cx a) { return cmk(a.x, -a.y); }
LQ_HD cx cscale(cx a, double s) { return cmk(a.x * s, a.y * s); }
LQ_HD cx cmul(cx a, cx b) { return cmk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
LQ_HD cx cmul_c(cx a, cx b) { return cmk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a * conj(b)
LQ_HD double cnorm2(cx a) { return a.x * a.x + a.y * a.y; }
// acc += a*b
LQ_HD void cfma(cx& acc, cx a, cx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// acc += a*conj(b)
LQ_HD void cfma_c(cx& acc, cx a, cx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.y, b.x, acc.y);
  acc.y = fma(-a.x, b.y, acc.y);
}
// acc += conj(a)*b
LQ_HD void cfma_ca(cx& acc, cx a, cx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}

// ---------------------------------------------------------------------------------------------- 3x3
struct M3 {
  cx e[9];  // row-major e[3*r + c]
};
LQ_HD M3 m3_zero() {
  M3 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = cmk(0.0, 0.0);
  return r;
}
LQ_HD M3 m3_ident() {
  M3 r = m3_zero();
  r.e[0].x = r.e[4].x = r.e[8].x = 1.0;
  return r;
}
LQ_HD M3 m3_add(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = cadd(a.e[k], b.e[k]);
  return r;
}
LQ_HD M3 m3_sub(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = csub(a.e[k], b.e[k]);
  return r;
}
LQ_HD M3 m3_adj(const M3& a) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.e[3 * i + j] = cconj(a.e[3 * j + i]);
  return r;
}
LQ_HD M3 m3_scale(const M3& a, double s) {
  M3 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = cscale(a.e[k], s);
  return r;
}
LQ_HD M3 m3_cscale(const M3& a, cx s) {
  M3 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = cmul(a.e[k], s);
  return r;
}
// acc += A*B and its two adjoint forms.  Every accumulator element sees the same sequence of FMAs as the plain
// triple loop over (i, j, k) with cfma / cfma_c / cfma_ca innermost (k ascending; within one k the real and imaginary
// updates in the order those helpers apply them), so the bits do not depend on the ordering chosen here.
// Device ordering ("operand-stationary"): the FMAs are issued in runs of six that share one multiplicand -- a scalar
// of A (nn, dn) or of B (nd) against the three columns / rows it meets.  A DFMA whose three register operands are all
// fresh issues every ~2.9 cycles on sm_100 instead of every 2 (measured: 24.7 against 35.2 TFLOP/s); one operand held
// over from the previous DFMA in the operand-reuse cache restores the full rate.  ptxas reorders plain C++ FMAs freely
// (its own order reuses an operand in ~40 % of the DFMAs, 28 TFLOP/s); `asm volatile` pins the order written here.
#if defined(__CUDA_ARCH__) && !defined(LQ_PLAIN_FMA_ORDER)
__device__ __forceinline__ double lq_vfma(double a, double b, double c) {
  double d;
  asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c));
  return d;
}
LQ_HD void m3_fma_nn(M3& acc, const M3& a, const M3& b) {  // acc_ij += a_ik b_kj
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double ax = a.e[3 * i + k].x, ay = a.e[3 * i + k].y;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        acc.e[3 * i + j].x = lq_vfma(ax, b.e[3 * k + j].x, acc.e[3 * i + j].x);
        acc.e[3 * i + j].y = lq_vfma(ax, b.e[3 * k + j].y, acc.e[3 * i + j].y);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        acc.e[3 * i + j].x = lq_vfma(-ay, b.e[3 * k + j].y, acc.e[3 * i + j].x);
        acc.e[3 * i + j].y = lq_vfma(ay, b.e[3 * k + j].x, acc.e[3 * i + j].y);
      }
    }
}
LQ_HD void m3_fma_nd(M3& acc, const M3& a, const M3& b) {  // acc_ij += a_ik conj(b_jk)
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double bx = b.e[3 * j + k].x, by = b.e[3 * j + k].y;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        acc.e[3 * i + j].x = lq_vfma(a.e[3 * i + k].x, bx, acc.e[3 * i + j].x);
        acc.e[3 * i + j].y = lq_vfma(a.e[3 * i + k].y, bx, acc.e[3 * i + j].y);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        acc.e[3 * i + j].x = lq_vfma(a.e[3 * i + k].y, by, acc.e[3 * i + j].x);
        acc.e[3 * i + j].y = lq_vfma(-a.e[3 * i + k].x, by, acc.e[3 * i + j].y);
      }
    }
}
LQ_HD void m3_fma_dn(M3& acc, const M3& a, const M3& b) {  // acc_ij += conj(a_ki) b_kj
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double ax = a.e[3 * k + i].x, ay = a.e[3 * k + i].y;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        acc.e[3 * i + j].x = lq_vfma(ax, b.e[3 * k + j].x, acc.e[3 * i + j].x);
        acc.e[3 * i + j].y = lq_vfma(ax, b.e[3 * k + j].y, acc.e[3 * i + j].y);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        acc.e[3 * i + j].x = lq_vfma(ay, b.e[3 * k + j].y, acc.e[3 * i + j].x);
        acc.e[3 * i + j].y = lq_vfma(-ay, b.e[3 * k + j].x, acc.e[3 * i + j].y);
      }
    }
}
#else
LQ_HD void m3_fma_nn(M3& acc, const M3& a, const M3& b) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) cfma(acc.e[3 * i + j], a.e[3 * i + k], b.e[3 * k + j]);
}
LQ_HD void m3_fma_nd(M3& acc, const M3& a, const M3& b) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) cfma_c(acc.e[3 * i + j], a.e[3 * i + k], b.e[3 * j + k]);
}
LQ_HD void m3_fma_dn(M3& acc, const M3& a, const M3& b) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) cfma_ca(acc.e[3 * i + j], a.e[3 * k + i], b.e[3 * k + j]);
}
#endif
// r = A*B etc.: the accumulators start from a zero ptxas cannot see through, so the first FMA of every element is an
// ordinary link of its chain and the operand-stationary order above survives scheduling (with the literal zero the
// eighteen chain heads have no predecessor and ptxas hoists them out of their runs)
#if !defined(LQ_HOST_EMU)
static __constant__ double lq_czero = 0.0;  // read through the constant bank: not foldable at compile time
#endif
LQ_HD M3 m3_zero_opaque() {
#if defined(__CUDA_ARCH__) && !defined(LQ_PLAIN_FMA_ORDER)
  const double z = lq_czero;
  M3 r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = cmk(z, z);
  return r;
#else
  return m3_zero();
#endif
}
LQ_HD M3 m3_mul_nn(const M3& a, const M3& b) {
  M3 r = m3_zero_opaque();
  m3_fma_nn(r, a, b);
  return r;
}
LQ_HD M3 m3_mul_nd(const M3& a, const M3& b) {
  M3 r = m3_zero_opaque();
  m3_fma_nd(r, a, b);
  return r;
}
LQ_HD M3 m3_mul_dn(const M3& a, const M3& b) {
  M3 r = m3_zero_opaque();
  m3_fma_dn(r, a, b);
  return r;
}
LQ_HD cx m3_trace(const M3& a) { return cmk(a.e[0].x + a.e[4].x + a.e[8].x, a.e[0].y + a.e[4].y + a.e[8].y); }
// Tr(A * B^dagger) = sum_rc A_rc conj(B_rc)
LQ_HD cx m3_trace_nd(const M3& a, const M3& b) {
  cx t = cmk(0.0, 0.0);
#pragma unroll
  for (int k = 0; k < 9; ++k) cfma_c(t, a.e[k], b.e[k]);
  return t;
}
LQ_HD cx m3_det(const M3& a) {
  cx minor1 = csub(cmul(a.e[4], a.e[8]), cmul(a.e[7], a.e[5]));
  cx minor2 = csub(cmul(a.e[3], a.e[8]), cmul(a.e[6], a.e[5]));
  cx minor3 = csub(cmul(a.e[3], a.e[7]), cmul(a.e[6], a.e[4]));
  return cadd(csub(cmul(a.e[0], minor1), cmul(a.e[1], minor2)), cmul(a.e[2], minor3));
}

// ---------------------------------------------------------------------------------------------- SoA access
LQ_HD M3 lq_load_link(const cx* LQ_RESTRICT U, const LqGeom& g, int dir, lq_i64 p) {
  M3 r;
  const cx* b = U + lq_addr(p, 9 * g.D, dir * 9);
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = LQ_LDG(b + k * 32);
  return r;
}
// plain (coherent) load: for kernels that also write the array they read
LQ_HD M3 lq_load_link_rw(const cx* U, const LqGeom& g, int dir, lq_i64 p) {
  M3 r;
  const cx* b = U + lq_addr(p, 9 * g.D, dir * 9);
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = b[k * 32];
  return r;
}
LQ_HD void lq_store_link(cx* U, const LqGeom& g, int dir, lq_i64 p, const M3& m) {
  cx* b = U + lq_addr(p, 9 * g.D, dir * 9);
#pragma unroll
  for (int k = 0; k < 9; ++k) b[k * 32] = m.e[k];
}
LQ_HD M3 lq_load_g(const cx* G, lq_i64 p) {
  M3 r;
  const cx* b = G + lq_addr(p, 9, 0);
#pragma unroll
  for (int k = 0; k < 9; ++k) r.e[k] = b[k * 32];
  return r;
}
LQ_HD void lq_store_g(cx* G, lq_i64 p, const M3& m) {
  cx* b = G + lq_addr(p, 9, 0);
#pragma unroll
  for (int k = 0; k < 9; ++k) b[k * 32] = m.e[k];
}
struct A8 {
  double e[8];
};
LQ_HD A8 lq_load_e(const cx* E, const LqGeom& g, int dir, lq_i64 p) {
  A8 r;
  const cx* b = E + lq_addr(p, 4 * g.D, dir * 4);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    cx v = b[q * 32];
    r.e[2 * q] = v.x;
    r.e[2 * q + 1] = v.y;
  }
  return r;
}
LQ_HD void lq_store_e(cx* E, const LqGeom& g, int dir, lq_i64 p, const A8& a) {
  cx* b = E + lq_addr(p, 4 * g.D, dir * 4);
#pragma unroll
  for (int q = 0; q < 4; ++q) b[q * 32] = cmk(a.e[2 * q], a.e[2 * q + 1]);
}

// ---------------------------------------------------------------------------------------------- su(3) algebra
// Su3Adjoint::to_matrix (field.rs:106-112) with the Gell-Mann/2 generators of su3.rs:24-194.
#define LQ_S3 0.288675134594812900  /* ONE_OVER_2_SQRT_3,   su3.rs:183 */
#define LQ_M3 -0.577350269189625800 /* MINUS_ONE_OVER_SQRT_3, su3.rs:182 */
LQ_HD M3 lq_adjoint_to_matrix(const A8& a) {
  M3 m;
  // explicit fma: the contraction must not depend on the kernel this is inlined into
  m.e[0] = cmk(fma(0.5, a.e[2], LQ_S3 * a.e[7]), 0.0);
  m.e[4] = cmk(fma(-0.5, a.e[2], LQ_S3 * a.e[7]), 0.0);
  m.e[8] = cmk(LQ_M3 * a.e[7], 0.0);
  m.e[1] = cmk(0.5 * a.e[0], -0.5 * a.e[1]);
  m.e[3] = cmk(0.5 * a.e[0], 0.5 * a.e[1]);
  m.e[2] = cmk(0.5 * a.e[3], -0.5 * a.e[4]);
  m.e[6] = cmk(0.5 * a.e[3], 0.5 * a.e[4]);
  m.e[5] = cmk(0.5 * a.e[5], -0.5 * a.e[6]);
  m.e[7] = cmk(0.5 * a.e[5], 0.5 * a.e[6]);
  return m;
}
// out_a = Tr(T_a W) for the 8 generators, as (re, im) pairs.
LQ_HD void lq_trace_gen(const M3& w, cx out[8]) {
  // T1: .5(W10 + W01); T2: .5 i (W01 - W10); T3: .5 (W00 - W11); T4: .5(W20+W02); T5: .5 i (W02 - W20)
  // T6: .5(W21+W12); T7: .5 i (W12 - W21); T8: S3 (W00 + W11) + M3 W22
  out[0] = cmk(0.5 * (w.e[3].x + w.e[1].x), 0.5 * (w.e[3].y + w.e[1].y));
  out[1] = cmk(-0.5 * (w.e[1].y - w.e[3].y), 0.5 * (w.e[1].x - w.e[3].x));
  out[2] = cmk(0.5 * (w.e[0].x - w.e[4].x), 0.5 * (w.e[0].y - w.e[4].y));
  out[3] = cmk(0.5 * (w.e[6].x + w.e[2].x), 0.5 * (w.e[6].y + w.e[2].y));
  out[4] = cmk(-0.5 * (w.e[2].y - w.e[6].y), 0.5 * (w.e[2].x - w.e[6].x));
  out[5] = cmk(0.5 * (w.e[7].x + w.e[5].x), 0.5 * (w.e[7].y + w.e[5].y));
  out[6] = cmk(-0.5 * (w.e[5].y - w.e[7].y), 0.5 * (w.e[5].x - w.e[7].x));
  out[7] = cmk(fma(LQ_S3, w.e[0].x + w.e[4].x, LQ_M3 * w.e[8].x), fma(LQ_S3, w.e[0].y + w.e[4].y, LQ_M3 * w.e[8].y));
}
// orthonormalize_matrix, su3.rs:279-303: Gram-Schmidt of columns 0,1; column 2 = conj(v1) x conj(v2).
// `try_normalize(eps)` leaves a vector untouched when its norm is <= eps.
LQ_HD M3 lq_orthonormalize(const M3& a) {
  cx v1[3] = {a.e[0], a.e[3], a.e[6]};
  cx v2[3] = {a.e[1], a.e[4], a.e[7]};
  double n1 = sqrt(cnorm2(v1[0]) + cnorm2(v1[1]) + cnorm2(v1[2]));
  if (n1 > LQ_EPS) {
#pragma unroll
    for (int k = 0; k < 3; ++k) v1[k] = cmk(v1[k].x / n1, v1[k].y / n1);
  }
  cx d = cmk(0.0, 0.0);
#pragma unroll
  for (int k = 0; k < 3; ++k) cfma_ca(d, v1[k], v2[k]);
  cx w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = csub(v2[k], cmul(v1[k], d));
  double n2 = sqrt(cnorm2(w[0]) + cnorm2(w[1]) + cnorm2(w[2]));
  if (n2 > LQ_EPS) {
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = cmk(w[k].x / n2, w[k].y / n2);
  }
  cx a1[3] = {cconj(v1[0]), cconj(v1[1]), cconj(v1[2])};
  cx b1[3] = {cconj(w[0]), cconj(w[1]), cconj(w[2])};
  cx cr[3] = {csub(cmul(a1[1], b1[2]), cmul(a1[2], b1[1])), csub(cmul(a1[2], b1[0]), cmul(a1[0], b1[2])),
              csub(cmul(a1[0], b1[1]), cmul(a1[1], b1[0]))};
  M3 r;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r.e[3 * k + 0] = v1[k];
    r.e[3 * k + 1] = w[k];
    r.e[3 * k + 2] = cr[k];
  }
  return r;
}
// su3_exp_i (su3.rs:820-855): exp(i sum_a e_a T_a), N = 26 Cayley-Hamilton recursion; t and d as field.rs:185-205.
LQ_HD M3 lq_su3_exp_i(const A8& a) {
  const double inv_fact[26] = {1.0,
                               1.0,
                               0.5,
                               1.0 / 6.0,
                               1.0 / 24.0,
                               1.0 / 120.0,
                               1.0 / 720.0,
                               1.0 / 5040.0,
                               1.0 / 40320.0,
                               1.0 / 362880.0,
                               1.0 / 3628800.0,
                               1.0 / 39916800.0,
                               1.0 / 479001600.0,
                               1.0 / 6227020800.0,
                               1.0 / 87178291200.0,
                               1.0 / 1307674368000.0,
                               1.0 / 20922789888000.0,
                               1.0 / 355687428096000.0,
                               1.0 / 6402373705728000.0,
                               1.0 / 121645100408832000.0,
                               1.0 / 2432902008176640000.0,
                               1.0 / 51090942171709440000.0,
                               1.0 / 1124000727777607680000.0,
                               1.0 / 25852016738884976640000.0,
                               1.0 / 620448401733239439360000.0,
                               1.0 / 15511210043330985984000000.0};
  M3 m = lq_adjoint_to_matrix(a);
  double ts = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) ts += a.e[k] * a.e[k];
  cx t = cmk(-0.5 * (ts / 2.0), 0.0);
  cx d = cmul(m3_det(m), cmk(0.0, 1.0));
  cx q0 = cmk(inv_fact[25], 0.0), q1 = cmk(0.0, 0.0), q2 = cmk(0.0, 0.0);
  for (int i = 24; i >= 0; --i) {
    cx q0n = cadd(cmk(inv_fact[i], 0.0), cmul(d, q2));
    cx q1n = cmul(cmk(0.0, 1.0), csub(q0, cmul(t, q2)));
    cx q2n = cmul(cmk(0.0, 1.0), q1);
    q0 = q0n;
    q1 = q1n;
    q2 = q2n;
  }
  M3 r = m3_add(m3_cscale(m, q1), m3_cscale(m3_mul_nn(m, m), q2));
  r.e[0] = cadd(r.e[0], q0);
  r.e[4] = cadd(r.e[4], q0);
  r.e[8] = cadd(r.e[8], q0);
  return r;
}
