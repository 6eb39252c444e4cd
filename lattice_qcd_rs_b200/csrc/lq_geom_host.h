// lq_geom_host.h -- host-side construction of LqGeom (shared by the C ABI and tools/kbench.cu).
#pragma once
#include <cstring>

#include "lq_common.cuh"

// returns 0, or -1 (LQ_E_BADARG) on an invalid lattice / process grid
static inline int init_geom(LqGeom& g, int D, const int64_t* gext, const int* nproc, const int* coord) {
  if (D < 2 || D > LQ_MAXD) return -1;
  memset(&g, 0, sizeof(g));
  g.D = D;
  lq_i64 ss = 1, gs = 1, ls = 1;
  for (int d = 0; d < LQ_MAXD; ++d) {
    if (d < D) {
      if (gext[d] < 2 || gext[d] > (1 << 20)) return -1;  // LatticeCyclic::new needs dim >= 2 (lattice.rs:190-201)
      int np = nproc ? nproc[d] : 1;
      if (np < 1 || gext[d] % np != 0) return -1;
      if (np > 1 && d < D - 2) return -1;  // only the two slowest directions may be split
      if (np > 1 && d == 0) return -1;
      g.gext[d] = (int)gext[d];
      g.ext[d] = (int)(gext[d] / np);
      if (np > 1 && g.ext[d] < 2) return -1;
      g.ghost[d] = np > 1 ? 1 : 0;
      g.goff[d] = np > 1 ? coord[d] * g.ext[d] : 0;
      if (np > 1 && (coord[d] < 0 || coord[d] >= np)) return -1;
    } else {
      g.gext[d] = g.ext[d] = 1;
      g.ghost[d] = 0;
      g.goff[d] = 0;
    }
    g.sext[d] = g.ext[d] + 2 * g.ghost[d];
    g.sstride[d] = ss;
    g.nstride[d] = ss;
    g.gstride[d] = gs;
    g.lstride[d] = ls;
    ss *= g.sext[d];
    gs *= g.gext[d];
    ls *= g.ext[d];
  }
  g.vol = ls;
  g.svol = ss;
  g.nchunk = (g.svol + 31) / 32;
  g.ne0 = (g.ext[0] + 1) / 2;
  return 0;
}


// Choose the brick the tuned kernels walk: `want[d]` is clipped to a divisor of ext[d]; tile[0] must be even,
// otherwise the whole row is used.  Returns the tile volume.
static inline int lq_set_tile(LqGeom& g, const int* want) {
  g.tvol = 1;
  for (int d = 0; d < LQ_MAXD; ++d) {
    int t = (d < g.D && want) ? want[d] : 1;
    if (t < 1) t = 1;
    if (t > g.ext[d]) t = g.ext[d];
    while (g.ext[d] % t) --t;
    if (d == 0 && (t & 1)) t = (g.ext[0] & 1) ? 0 : g.ext[0];
    g.tile[d] = t;
    g.ntile[d] = t ? g.ext[d] / t : 0;
    g.tvol *= t;
  }
  return g.tvol;  // 0: odd x0 extent, tiled walk unavailable (callers use the row walk)
}
