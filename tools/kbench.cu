// tools/kbench.cu -- micro-benchmark of the MD hot-loop kernel variants on one GPU (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I lattice_qcd_rs_b200/csrc tools/kbench.cu -o tools/kbench
//   tools/kbench [extent=32] [reps=10]
// Every variant is checked against the generic functor kernel (KEfieldLinkStep) before it is timed.
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

#include "lq_geom_host.h"
#include "lq_md_variants.cuh"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

template <class F>
__global__ void __launch_bounds__(128) gen_k(F f, lq_i64 n) {
  lq_i64 i = (lq_i64)blockIdx.x * 128 + threadIdx.x;
  if (i < n) f(i);
}
template <class F>
static void gen_launch(lq_i64 n, const F& f) {
  gen_k<F><<<(unsigned)((n + 127) / 128), 128>>>(f, n);
  CK(cudaGetLastError());
}

// FP64 FMA peak: 8 independent chains per thread
__global__ void dfma_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// FP64 rate of the 3x3 complex matrix-product code itself, operands resident in registers (no memory traffic): the
// staple recurrence t = a b^+, acc += t c^+ of the MD kernel, REP times per thread.  MINB sets the register budget
// (3 -> 168 registers, 12 warps/SM as lq_md4_kernel).
template <int MINB>
__global__ void __launch_bounds__(128, MINB) mm_peak(const cx* __restrict__ in, cx* out, int reps) {
  M3 a, b, c, acc = m3_zero();
  const int t0 = blockIdx.x * 128 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    a.e[k] = in[(t0 * 27 + k) & 0xfffff];
    b.e[k] = in[(t0 * 27 + 9 + k) & 0xfffff];
    c.e[k] = in[(t0 * 27 + 18 + k) & 0xfffff];
  }
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
    M3 t = m3_mul_nd(a, b);
    m3_fma_nd(acc, t, c);
    M3 t2 = m3_mul_nn(c, a);
    m3_fma_dn(acc, t2, b);
#pragma unroll
    for (int k = 0; k < 9; ++k) a.e[k].x += 1e-9 * acc.e[8 - k].y;  // keep the products loop-carried
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) out[t0 * 9 + k] = acc.e[k];
}
// streaming copy (HBM peak on this box, double2)
__global__ void copy_k(const double2* __restrict__ a, double2* __restrict__ b, lq_i64 n) {
  lq_i64 i = (lq_i64)blockIdx.x * blockDim.x + threadIdx.x;
  lq_i64 stride = (lq_i64)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}

// L2 -> SM read bandwidth: every block streams the same 48 MiB window (L2 resident) with L1 bypass (ld.global.cg)
__global__ void l2_read_k(const double2* __restrict__ a, lq_i64 n, int passes, double* out) {
  double acc = 0.0;
  for (int ps = 0; ps < passes; ++ps) {
    lq_i64 i = ((lq_i64)blockIdx.x * 7919 * blockDim.x + threadIdx.x) % n;
    for (lq_i64 k = 0; k < n / ((lq_i64)gridDim.x * blockDim.x) * 8; ++k) {
      double2 v = __ldcg(a + i);
      acc += v.x + v.y;
      i += (lq_i64)blockDim.x * gridDim.x;
      if (i >= n) i -= n;
    }
  }
  if (acc == 1.2345e-300) out[0] = acc;
}

struct Variant {
  std::string name;
  std::function<void()> run;
  const void* func;
  int block;
};

int main(int argc, char** argv) {
  int L = argc > 1 ? atoi(argv[1]) : 32;
  int reps = argc > 2 ? atoi(argv[2]) : 10;
  const char* filter = argc > 3 ? argv[3] : "";  // run only the variants whose name contains this
  int64_t ext[4] = {L, L, L, L};
  LqGeom g;
  if (init_geom(g, 4, ext, nullptr, nullptr)) return 1;
  int rowtile[4] = {L, 1, 1, 1};
  lq_set_tile(g, rowtile);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, L=%d, vol=%lld, links=%lld\n", prop.name, prop.multiProcessorCount, L, g.vol, g.vol * 4);

  size_t ub = (size_t)g.nchunk * 32 * 36 * sizeof(cx), eb = (size_t)g.nchunk * 32 * 16 * sizeof(cx);
  cx *U, *U2, *E, *E0, *Uref, *Eref;
  CK(cudaMalloc(&U, ub)); CK(cudaMalloc(&U2, ub)); CK(cudaMalloc(&Uref, ub));
  CK(cudaMalloc(&E, eb)); CK(cudaMalloc(&E0, eb)); CK(cudaMalloc(&Eref, eb));
  CK(cudaMemset(U, 0, ub)); CK(cudaMemset(U2, 0, ub)); CK(cudaMemset(E, 0, eb));
  gen_launch(lq_link_items(g), KLinksRandom<4>{g, U, 0x1234567ull, 0});
  gen_launch(lq_link_items(g), KMomentaRefresh<4>{g, E0, 0x1234567ull, 1, 0.1});
  CK(cudaDeviceSynchronize());
  const double coef = -sqrt(2.0 / 3.0), c_u = sqrt(6.0), dt = 0.01;
  const lq_i64 nl = g.vol * 4;

  // ---- peaks on this box
  {
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 256));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int iters = 200000;
    dfma_peak<<<148 * 8, 256>>>(out, 1000);
    CK(cudaEventRecord(e0));
    dfma_peak<<<148 * 8, 256>>>(out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 8 * iters * 148.0 * 8 * 256;
    printf("FP64 FMA peak: %.2f TFLOP/s (%.3f ms)\n", fl / ms / 1e9, ms);
    {
      cx* mo;
      CK(cudaMalloc(&mo, sizeof(cx) * 148 * 24 * 128 * 9));
      auto run_mm = [&](auto kern, const char* name, int blocks) {
        const int reps = 2000;
        kern<<<blocks, 128>>>(U, mo, 10);
        CK(cudaEventRecord(e0));
        kern<<<blocks, 128>>>(U, mo, reps);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        // 4 products x 27 complex FMAs x 4 DFMA x 2 flop
        printf("3x3 complex matmul code in registers, %s: %.2f TFLOP/s\n", name, 4.0 * 27 * 8 * reps * blocks * 128.0 / ms / 1e9);
      };
      run_mm(mm_peak<3>, "168-register budget, 12 warps/SM", 148 * 3 * 8);
      run_mm(mm_peak<2>, "255-register budget, 8 warps/SM", 148 * 2 * 8);
      run_mm(mm_peak<4>, "128-register budget, 16 warps/SM", 148 * 4 * 8);
      CK(cudaFree(mo));
    }
    lq_i64 n = (lq_i64)g.nchunk * 32 * 36;
    copy_k<<<148 * 16, 256>>>(U, U2, n);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 5; ++r) copy_k<<<148 * 16, 256>>>(U, U2, n);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("copy (read+write) : %.1f GB/s\n", 5.0 * 2 * n * 16 / ms / 1e6);
    {
      lq_i64 n2 = (lq_i64)48 * 1024 * 1024 / 16;
      int blocks = 148 * 8, threads = 256, passes = 4;
      l2_read_k<<<blocks, threads>>>(U, n2, 1, out);
      CK(cudaEventRecord(e0));
      l2_read_k<<<blocks, threads>>>(U, n2, passes, out);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      double bytes = (double)passes * (n2 / ((lq_i64)blocks * threads) * 8) * blocks * threads * 16.0;
      printf("L2->SM read (48 MiB window, ld.cg): %.1f GB/s\n", bytes / ms / 1e6);
    }
    CK(cudaFree(out));
  }

  // ---- reference output from the generic functor kernel
  CK(cudaMemcpy(E, E0, eb, cudaMemcpyDeviceToDevice));
  gen_launch(lq_link_items(g), KEfieldLinkStep<4>{g, U, Uref, E, coef, dt / 2, dt, c_u, 2, 0});
  CK(cudaMemcpy(Eref, E, eb, cudaMemcpyDeviceToDevice));
  CK(cudaDeviceSynchronize());
  std::vector<double> hUref(ub / 8), hEref(eb / 8), hU(ub / 8), hE(eb / 8);
  CK(cudaMemcpy(hUref.data(), Uref, ub, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hEref.data(), Eref, eb, cudaMemcpyDeviceToHost));

  std::vector<Variant> vs;
  auto tiled = [&](int a, int b, int c, int d) {
    LqGeom t = g;
    int w[4] = {a, b, c, d};
    lq_set_tile(t, w);
    return t;
  };
  vs.push_back({"generic KEfieldLinkStep (block 128)", [&] { gen_launch(lq_link_items(g), KEfieldLinkStep<4>{g, U, U2, E, coef, dt / 2, dt, c_u, 2, 0}); },
                (const void*)gen_k<KEfieldLinkStep<4>>, 128});
#define V1(BLOCK, MINB, MAP, GEOM, LABEL)                                                                          \
  vs.push_back({std::string("v1 block=" #BLOCK " minb=" #MINB " ") + LABEL,                                         \
                [&, gg = GEOM] {                                                                                    \
                  lq_md_link_kernel<BLOCK, MINB, MAP, 1><<<(unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4)), BLOCK>>>( \
                      gg, U, U2, E, coef, dt / 2, dt, c_u, 2);                                                      \
                  CK(cudaGetLastError());                                                                           \
                },                                                                                                  \
                (const void*)lq_md_link_kernel<BLOCK, MINB, MAP, 1>, BLOCK})
  V1(128, 1, 0, g, "row");
  V1(128, 2, 0, g, "row");
  V1(128, 3, 0, g, "row");
  V1(128, 4, 0, g, "row");
  V1(256, 1, 0, g, "row");
  V1(256, 2, 0, g, "row");
  V1(64, 4, 0, g, "row");
  V1(64, 8, 0, g, "row");
  V1(128, 3, 1, tiled(16, 2, 1, 1), "tile16x2x1x1");
  V1(128, 3, 1, tiled(8, 2, 2, 1), "tile8x2x2x1");
  V1(128, 3, 1, tiled(8, 4, 1, 1), "tile8x4x1x1");
  V1(256, 2, 1, tiled(16, 2, 2, 1), "tile16x2x2x1");
  V1(256, 2, 1, tiled(8, 2, 2, 2), "tile8x2x2x2");
  V1(256, 2, 1, tiled(16, 4, 1, 1), "tile16x4x1x1");
  V1(256, 2, 1, tiled(32, 2, 1, 1), "tile32x2x1x1");
  V1(512, 1, 1, tiled(8, 4, 2, 2), "tile8x4x2x2");
  V1(512, 1, 1, tiled(16, 2, 2, 2), "tile16x2x2x2");
  V1(512, 1, 1, tiled(32, 2, 2, 1), "tile32x2x2x1");
  {
    LqGeom fake = g;  // neighbours in x1..x3 collapse onto the site itself: perfect-locality bound of this code
    fake.nstride[1] = fake.nstride[2] = fake.nstride[3] = 0;
    V1(128, 3, 0, fake, "row FAKE-LOCALITY (not a valid result)");
    V1(256, 2, 0, fake, "row FAKE-LOCALITY (not a valid result)");
  }
#define V3(BLOCK, MINB, MAP, GEOM, LABEL)                                                                          \
  vs.push_back({std::string("v3 nu-loop block=" #BLOCK " minb=" #MINB " ") + LABEL,                                 \
                [&, gg = GEOM] {                                                                                    \
                  lq_md_link_loop_kernel<BLOCK, MINB, MAP, 1><<<(unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4)), BLOCK>>>( \
                      gg, U, U2, E, coef, dt / 2, dt, c_u, 2);                                                      \
                  CK(cudaGetLastError());                                                                           \
                },                                                                                                  \
                (const void*)lq_md_link_loop_kernel<BLOCK, MINB, MAP, 1>, BLOCK})
  V3(128, 1, 0, g, "row");
  V3(128, 3, 0, g, "row");
  V3(128, 4, 0, g, "row");
  V3(256, 1, 0, g, "row");
  V3(256, 2, 0, g, "row");
  V3(512, 1, 1, tiled(32, 2, 2, 1), "tile32x2x2x1");
#define V4(BLOCK, MINB)                                                                                            \
  vs.push_back({std::string("v4 lean nu-loop block=" #BLOCK " minb=" #MINB),                                        \
                [&] {                                                                                               \
                  lq_md4x_kernel<BLOCK, MINB, 1><<<(unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4)), BLOCK>>>(      \
                      g, U, U2, E, coef, dt / 2, dt, c_u, 2, nullptr, 0);                                             \
                  CK(cudaGetLastError());                                                                           \
                },                                                                                                  \
                (const void*)lq_md4x_kernel<BLOCK, MINB, 1>, BLOCK})
  V4(128, 3);
  V4(64, 6);
  V4(64, 7);
  V4(192, 2);
#define V4F(BLOCK, MINB, FLAGS, CARVE)                                                                             \
  vs.push_back({std::string("v4 block=" #BLOCK " minb=" #MINB " flags=" #FLAGS " carveout=" #CARVE),              \
                [&] {                                                                                               \
                  if (CARVE >= 0)                                                                                   \
                    cudaFuncSetAttribute(lq_md4x_kernel<BLOCK, MINB, 1, FLAGS>,                                      \
                                         cudaFuncAttributePreferredSharedMemoryCarveout, CARVE);                   \
                  lq_md4x_kernel<BLOCK, MINB, 1, FLAGS><<<(unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4)), BLOCK>>>( \
                      g, U, U2, E, coef, dt / 2, dt, c_u, 2, nullptr, 0);                                             \
                  CK(cudaGetLastError());                                                                           \
                },                                                                                                  \
                (const void*)lq_md4x_kernel<BLOCK, MINB, 1, FLAGS>, BLOCK})
#define V4P(BLOCK, MINB, FLAGS, GRID)                                                                              \
  vs.push_back({std::string("v4 persistent block=" #BLOCK " minb=" #MINB " flags=" #FLAGS " grid=" #GRID),        \
                [&] {                                                                                               \
                  lq_md4x_kernel<BLOCK, MINB, 1, FLAGS><<<GRID, BLOCK>>>(g, U, U2, E, coef, dt / 2, dt, c_u, 2,     \
                                                                       nullptr, 0);                                \
                  CK(cudaGetLastError());                                                                           \
                },                                                                                                  \
                (const void*)lq_md4x_kernel<BLOCK, MINB, 1, FLAGS>, BLOCK})
  V4P(128, 3, 18, 444);
  V4P(128, 3, 18, 888);
  V4P(128, 3, 18, 1776);
  V4P(128, 3, 18, 3552);
  V4P(128, 3, 18, 8192);
  V4P(256, 1, 18, 148);
  V4P(256, 1, 18, 592);
  V4F(128, 3, 2, -1);
  V4F(128, 3, 66, -1);
  V4F(128, 2, 66, -1);
  V4F(128, 1, 66, -1);
  V4F(256, 1, 66, -1);
  V4F(128, 3, 8, -1);
  V4F(128, 3, 10, -1);
  V4F(128, 3, 4, -1);
  V4F(128, 3, 34, -1);
  V4F(128, 3, 32, -1);
  V4F(128, 3, 130, -1);
  V4F(128, 3, 130, 0);
  V4F(128, 2, 130, -1);
#define V5(BLOCK, MINB, PIPE)                                                                                      \
  vs.push_back({std::string("v5 pipelined block=" #BLOCK " minb=" #MINB " pipe=" #PIPE),                            \
                [&] {                                                                                               \
                  lq_md5_kernel<BLOCK, MINB, 1, PIPE><<<(unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4)), BLOCK>>>( \
                      g, U, U2, E, coef, dt / 2, dt, c_u, 2);                                                       \
                  CK(cudaGetLastError());                                                                           \
                },                                                                                                  \
                (const void*)lq_md5_kernel<BLOCK, MINB, 1, PIPE>, BLOCK})
  V5(128, 3, 1);
  V5(128, 3, 3);
  V5(128, 2, 1);
  V5(128, 2, 3);
  V5(64, 6, 1);
  V5(64, 6, 3);
  V5(64, 4, 3);
  V5(256, 1, 3);
#define V6(BLOCK, MINB)                                                                                            \
  vs.push_back({std::string("v6 product-pipelined block=" #BLOCK " minb=" #MINB),                                   \
                [&] {                                                                                               \
                  lq_md6_kernel<BLOCK, MINB, 1><<<(unsigned)((g.vol + BLOCK / 4 - 1) / (BLOCK / 4)), BLOCK>>>(      \
                      g, U, U2, E, coef, dt / 2, dt, c_u, 2);                                                       \
                  CK(cudaGetLastError());                                                                           \
                },                                                                                                  \
                (const void*)lq_md6_kernel<BLOCK, MINB, 1>, BLOCK})
  V6(128, 3);
  V6(128, 4);
  V6(128, 5);
  V6(64, 8);
  V6(64, 10);
  V6(256, 2);
#define V2(MINB, MAP, GEOM, LABEL)                                                                           \
  vs.push_back({std::string("v2 nu-split block=384 minb=" #MINB " ") + LABEL,                                 \
                [&, gg = GEOM] {                                                                              \
                  lq_md_nusplit_kernel<MINB, MAP, 1><<<(unsigned)((g.vol + 31) / 32), 384>>>(gg, U, U2, E, coef, \
                                                                                             dt / 2, dt, c_u, 2); \
                  CK(cudaGetLastError());                                                                     \
                },                                                                                            \
                (const void*)lq_md_nusplit_kernel<MINB, MAP, 1>, 384})
  V2(1, 0, g, "row");
  V2(2, 0, g, "row");
  V2(1, 1, tiled(16, 2, 1, 1), "tile16x2x1x1");
  V2(1, 1, tiled(8, 2, 2, 1), "tile8x2x2x1");

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("%-52s %5s %4s %9s %9s %8s %10s %10s\n", "variant", "regs", "occ", "ms", "GB/s(alg)", "TF/s", "errE", "errU");
  for (auto& v : vs) {
    if (filter[0] && v.name.find(filter) == std::string::npos && v.name.find("generic") == std::string::npos) continue;
    cudaFuncAttributes at;
    CK(cudaFuncGetAttributes(&at, v.func));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.func, v.block, 0));
    CK(cudaMemcpy(E, E0, eb, cudaMemcpyDeviceToDevice));
    CK(cudaMemset(U2, 0, ub));
    v.run();
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hU.data(), U2, ub, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hE.data(), E, eb, cudaMemcpyDeviceToHost));
    double errE = 0, errU = 0;
    for (size_t i = 0; i < hE.size(); ++i) errE = fmax(errE, fabs(hE[i] - hEref[i]));
    for (size_t i = 0; i < hU.size(); ++i) errU = fmax(errU, fabs(hU[i] - hUref[i]));
    CK(cudaMemcpy(E, E0, eb, cudaMemcpyDeviceToDevice));
    v.run();
    v.run();
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) v.run();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    printf("%-52s %5d %4d %9.4f %9.1f %8.2f %10.2e %10.2e\n", v.name.c_str(), at.numRegs, occ * v.block / 32, ms,
           416.0 * nl / ms / 1e6, 3148.0 * nl / ms / 1e9, errE, errU);
    fflush(stdout);
  }
  // ---- checkerboard sweep sub-steps (8 launches = one sweep); every variant must reproduce the bits of PF = 0
  {
    cx* W;
    CK(cudaMalloc(&W, ub));
    std::vector<double> hW0(ub / 8), hW(ub / 8);
    struct SV {
      std::string name;
      std::function<void(int, int)> run;
      const void* func;
      int block;
    };
    std::vector<SV> sv;
#define SW(BLOCK, MINB, KIND)                                                                                  \
  sv.push_back({std::string(KIND == 0 ? "heatbath" : "overrelax") + " block=" #BLOCK " minb=" #MINB,    \
                [&](int mu, int par) {                                                                              \
                  lq_sweep4_kernel<BLOCK, MINB, KIND><<<(unsigned)((g.vol / 2 + BLOCK - 1) / BLOCK), BLOCK>>>(  \
                      g, W, mu, par, 0, 0, 6.0, 0x777ull, 5ull);                                                    \
                  CK(cudaGetLastError());                                                                           \
                },                                                                                                  \
                (const void*)lq_sweep4_kernel<BLOCK, MINB, KIND>, BLOCK})
    SW(128, 3, 0);
    SW(128, 4, 0);
    SW(128, 5, 0);
    SW(128, 6, 0);
    SW(64, 8, 0);
    SW(64, 10, 0);
    SW(256, 2, 0);
    SW(128, 3, 1);
    SW(128, 4, 1);
    SW(128, 5, 1);
    SW(128, 6, 1);
    printf("%-52s %5s %4s %9s %9s %10s\n", "sweep variant (8 sub-steps)", "regs", "occ", "ms/sweep", "GB/s(alg)", "maxdiff");
    for (size_t vi = 0; vi < sv.size(); ++vi) {
      auto& v = sv[vi];
      if (filter[0] && std::string("sweep").find(filter) == std::string::npos && v.name.find(filter) == std::string::npos) continue;
      cudaFuncAttributes at;
      CK(cudaFuncGetAttributes(&at, v.func));
      int occ = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.func, v.block, 0));
      CK(cudaMemcpy(W, U, ub, cudaMemcpyDeviceToDevice));
      for (int mu = 0; mu < 4; ++mu)
        for (int par = 0; par < 2; ++par) v.run(mu, par);
      CK(cudaDeviceSynchronize());
      const bool first_of_kind = v.name.find("block=128 minb=3") != std::string::npos;
      CK(cudaMemcpy((first_of_kind ? hW0 : hW).data(), W, ub, cudaMemcpyDeviceToHost));
      double diff = 0;
      if (!first_of_kind)
        for (size_t i = 0; i < hW.size(); ++i) diff = fmax(diff, fabs(hW[i] - hW0[i]));
      CK(cudaEventRecord(e0));
      for (int r = 0; r < reps; ++r)
        for (int mu = 0; mu < 4; ++mu)
          for (int par = 0; par < 2; ++par) v.run(mu, par);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ms /= reps;
      // algorithmic bytes of a sweep: 8 sub-steps x (all links read once + 1/8 written) = 1296 B per link update
      printf("%-52s %5d %4d %9.4f %9.1f %10.2e\n", v.name.c_str(), at.numRegs, occ * v.block / 32, ms,
             1296.0 * nl / ms / 1e6, diff);
      fflush(stdout);
    }
    CK(cudaFree(W));
  }
  return 0;
}
