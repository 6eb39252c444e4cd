// tools/kbench2.cu -- round-2 micro-benchmarks (development tool, one GPU):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DLQ_TUNED_NO_LAUNCHERS -I lattice_qcd_rs_b200/csrc -I tools tools/kbench2.cu -o tools/kbench2
//   tools/kbench2 [extent=32] [reps=10] [filter]
// (1) DFMA-stream efficiency of the 3x3 complex product code against occupancy (operands resident in registers);
// (2) the fused MD kernel: product kernel lq_md4_kernel, V7 (straight-line product pipeline, 128 / 96 registers),
//     V8 (row split).  Every variant is compared with the generic functor KEfieldLinkStep before it is timed.
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
__global__ void dfma_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// DFMA with three distinct register operands per instruction (8 accumulators x 4 multiplicand pairs)
__global__ void dfma_peak3(double* out, const double* in, int iters) {
  double a[8], x[4], y[4];
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-9 + k;
  for (int k = 0; k < 4; ++k) {
    x[k] = in[threadIdx.x & 63] + k;
    y[k] = in[(threadIdx.x + 7) & 63] * 1e-7 + k * 1e-8;
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(x[k & 3], y[(k + (k >> 2)) & 3], a[k]);
  }
  double s = 0;
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the staple recurrence t = a b^+, acc += t c^+ with a, b, c resident (90 registers live): DFMA-stream rate of the product
// code at 3 / 4 / 5 warps per scheduler
template <int MINB>
__global__ void __launch_bounds__(128, MINB) mm_stream(const cx* __restrict__ in, cx* out, int reps) {
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
#pragma unroll
    for (int k = 0; k < 9; ++k) a.e[k].x += 1e-9 * acc.e[8 - k].y;  // keep the products loop-carried
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) out[t0 * 9 + k] = acc.e[k];
}

// two ACCUMULATING products per iteration, each fenced into its own basic block (one-trip loop): the operand-stationary
// order of lq_common.cuh survives ptxas (runs of six DFMAs sharing a multiplicand) -- the ceiling of that technique
template <int MINB>
__global__ void __launch_bounds__(128, MINB) mm_chain(const cx* __restrict__ in, cx* out, int reps, int one) {
  M3 a, b, c, acc = m3_zero(), acc2 = m3_zero();
  const int t0 = blockIdx.x * 128 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    a.e[k] = in[(t0 * 27 + k) & 0xfffff];
    b.e[k] = in[(t0 * 27 + 9 + k) & 0xfffff];
    c.e[k] = in[(t0 * 27 + 18 + k) & 0xfffff];
  }
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int f = 0; f < one; ++f) m3_fma_nd(acc, a, b);
#pragma unroll 1
    for (int f = 0; f < one; ++f) m3_fma_nn(acc2, acc, c);
#pragma unroll
    for (int k = 0; k < 9; ++k) a.e[k].x += 1e-9 * acc2.e[8 - k].y;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) out[t0 * 9 + k] = cadd(acc.e[k], acc2.e[k]);
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
  const char* filter = argc > 3 ? argv[3] : "";
  int64_t ext[4] = {L, L, L, L};
  LqGeom g;
  if (init_geom(g, 4, ext, nullptr, nullptr)) return 1;
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
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  {
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 256));
    int iters = 100000;
    dfma_peak<<<148 * 8, 256>>>(out, 1000);
    CK(cudaEventRecord(e0));
    dfma_peak<<<148 * 8, 256>>>(out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("FP64 FMA peak (1 varying operand): %.2f TFLOP/s\n", 2.0 * 8 * iters * 148.0 * 8 * 256 / ms / 1e9);
    for (int bl : {3, 4, 6, 8, 16}) {
      dfma_peak3<<<148 * bl, 128>>>(out, (const double*)U, 1000);
      CK(cudaEventRecord(e0));
      dfma_peak3<<<148 * bl, 128>>>(out, (const double*)U, iters);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("FP64 FMA, 3 distinct register operands, %2d warps/SM: %.2f TFLOP/s\n", bl * 4,
             2.0 * 8 * iters * 148.0 * bl * 128 / ms / 1e9);
    }
    cx* mo;
    CK(cudaMalloc(&mo, sizeof(cx) * 148 * 64 * 128 * 9));
    auto run_mm = [&](auto kern, const char* name, int blocks) {
      const int r = 4000;
      cudaFuncAttributes at;
      CK(cudaFuncGetAttributes(&at, (const void*)kern));
      kern<<<blocks, 128>>>(U, mo, 10);
      CK(cudaEventRecord(e0));
      kern<<<blocks, 128>>>(U, mo, r);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("3x3 complex product stream in registers, %s (%d regs, %zu B local): %.2f TFLOP/s\n", name, at.numRegs,
             at.localSizeBytes, 2.0 * 27 * 8 * r * blocks * 128.0 / ms / 1e9);
    };
    run_mm(mm_stream<2>, " 8 warps/SM", 148 * 2 * 8);
    run_mm(mm_stream<3>, "12 warps/SM", 148 * 3 * 8);
    run_mm(mm_stream<4>, "16 warps/SM", 148 * 4 * 8);
    run_mm(mm_stream<5>, "20 warps/SM", 148 * 5 * 8);
    auto run_ch = [&](auto kern, const char* name, int blocks) {
      const int r = 4000;
      cudaFuncAttributes at;
      CK(cudaFuncGetAttributes(&at, (const void*)kern));
      kern<<<blocks, 128>>>(U, mo, 10, 1);
      CK(cudaEventRecord(e0));
      kern<<<blocks, 128>>>(U, mo, r, 1);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("fenced accumulating products, operand-stationary order, %s (%d regs, %zu B local): %.2f TFLOP/s\n", name,
             at.numRegs, at.localSizeBytes, 2.0 * 27 * 8 * r * blocks * 128.0 / ms / 1e9);
    };
    run_ch(mm_chain<2>, " 8 warps/SM", 148 * 2 * 8);
    run_ch(mm_chain<3>, "12 warps/SM", 148 * 3 * 8);
    run_ch(mm_chain<4>, "16 warps/SM", 148 * 4 * 8);
    run_ch(mm_chain<6>, "24 warps/SM", 148 * 6 * 8);
    CK(cudaFree(mo));
    CK(cudaFree(out));
  }
  CK(cudaMemcpy(E, E0, eb, cudaMemcpyDeviceToDevice));
  gen_launch(lq_link_items(g), KEfieldLinkStep<4>{g, U, Uref, E, coef, dt / 2, dt, c_u, 2, 0});
  CK(cudaMemcpy(Eref, E, eb, cudaMemcpyDeviceToDevice));
  CK(cudaDeviceSynchronize());
  std::vector<double> hUref(ub / 8), hEref(eb / 8), hU(ub / 8), hE(eb / 8);
  CK(cudaMemcpy(hUref.data(), Uref, ub, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hEref.data(), Eref, eb, cudaMemcpyDeviceToHost));

  std::vector<Variant> vs;
  const unsigned nb128 = (unsigned)((g.vol + 31) / 32);
#define ADD(NAME, KERN, GRID, BLOCK, ...)                                                  \
  vs.push_back({NAME, [&] { KERN<<<GRID, BLOCK>>>(__VA_ARGS__); CK(cudaGetLastError()); }, \
                (const void*)KERN, BLOCK})
  ADD("md4 product kernel <128,3>", (lq_md4_kernel<128, 3, 1>), nb128, 128, g, U, U2, E, coef, dt / 2, dt, c_u, 2, nullptr, 0, LqFold{});
  ADD("v7 straight-line pipeline <128,3> 168r", (lq_md7_kernel<128, 3, 1>), nb128, 128, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v7 straight-line pipeline <128,4> 128r", (lq_md7_kernel<128, 4, 1>), nb128, 128, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v7 straight-line pipeline <128,5>  96r", (lq_md7_kernel<128, 5, 1>), nb128, 128, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v7 straight-line pipeline <256,2> 128r", (lq_md7_kernel<256, 2, 1>), (nb128 + 1) / 2, 256, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v7 straight-line pipeline <64,8>  128r", (lq_md7_kernel<64, 8, 1>), nb128 * 2, 64, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v7 straight-line pipeline <128,2> 255r", (lq_md7_kernel<128, 2, 1>), nb128, 128, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v7 straight-line pipeline <64,6>  168r", (lq_md7_kernel<64, 6, 1>), nb128 * 2, 64, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v7 straight-line pipeline <192,2> 168r", (lq_md7_kernel<192, 2, 1>), (unsigned)((g.vol + 47) / 48), 192, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v10 two staples ahead <128,3> nu+1..", (lq_md10_kernel<128, 3, 1, 0>), nb128, 128, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v10 two staples ahead <128,3> nu asc", (lq_md10_kernel<128, 3, 1, 1>), nb128, 128, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v10 two staples ahead <128,2> 255r", (lq_md10_kernel<128, 2, 1, 0>), nb128, 128, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v9 fenced products <128,3> 168r", (lq_md9_kernel<128, 3, 1>), nb128, 128, g, U, U2, E, coef, dt / 2, dt, c_u, 2, 1);
  ADD("v9 fenced products <128,4> 128r", (lq_md9_kernel<128, 4, 1>), nb128, 128, g, U, U2, E, coef, dt / 2, dt, c_u, 2, 1);
  ADD("v9 fenced products <256,1> 255r", (lq_md9_kernel<256, 1, 1>), (nb128 + 1) / 2, 256, g, U, U2, E, coef, dt / 2, dt, c_u, 2, 1);
  // V11: TMA-staged operands (persistent, one block per SM)
  if (g.ext[0] == 32) {
    int nsm11 = 148;
    CK(cudaDeviceGetAttribute(&nsm11, cudaDevAttrMultiProcessorCount, 0));
    const int ntiles11 = (int)(g.vol / 32);
    CK(cudaFuncSetAttribute(lq_md11_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LQ11_SMEM));
    vs.push_back({"v11 TMA-staged operands, 8+1 warps, 1 block/SM", [=] {
                    lq_md11_tma_kernel<1><<<nsm11, 288, LQ11_SMEM>>>(g, U, U2, E, coef, dt / 2, dt, c_u, 2, ntiles11);
                    CK(cudaGetLastError());
                  },
                  (const void*)lq_md11_tma_kernel<1>, 288});
  }
  ADD("v8 row split <384,1>", (lq_md8_rowsplit_kernel<1, 1>), nb128, 384, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v8 row split <384,2>", (lq_md8_rowsplit_kernel<2, 1>), nb128, 384, g, U, U2, E, coef, dt / 2, dt, c_u, 2);
  ADD("v8 row split <384,3>", (lq_md8_rowsplit_kernel<3, 1>), nb128, 384, g, U, U2, E, coef, dt / 2, dt, c_u, 2);

  printf("%-46s %5s %6s %4s %9s %9s %8s %10s %10s\n", "variant", "regs", "local", "w/SM", "ms", "GB/s(alg)", "TF/s", "errE", "errU");
  for (auto& v : vs) {
    if (filter[0] && v.name.find(filter) == std::string::npos) continue;
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
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    printf("%-46s %5d %6zu %4d %9.4f %9.1f %8.2f %10.2e %10.2e\n", v.name.c_str(), at.numRegs, at.localSizeBytes,
           occ * v.block / 32, ms, 416.0 * nl / ms / 1e6, 3148.0 * nl / ms / 1e9, errE, errU);
    fflush(stdout);
  }
  // ---- checkerboard sweep sub-steps (8 launches = one sweep)
  {
    cx* W;
    CK(cudaMalloc(&W, ub));
    struct SV {
      std::string name;
      std::function<void(int, int)> run;
      const void* func;
      int block;
      int family;  // 0 / 2 / 4: heat bath / SVD over-relaxation / SU(2) over-relaxation (+1: compared with the first)
      int smem = 0;
    };
    std::vector<SV> sv;
    const unsigned nbs = (unsigned)((g.vol / 2 + 127) / 128);
#define SWV(NAME, FAM, KERN, ...)                                                                    \
  sv.push_back({NAME, [&](int mu, int par) { KERN<<<nbs, 128>>>(__VA_ARGS__); CK(cudaGetLastError()); }, \
                (const void*)KERN, 128, FAM})
    SWV("heatbath  product kernel (pipelined staples)", 0, (lq_sweep4_kernel<128, 3, 0>), g, W, mu, par, 0, 0, 6.0, 0x777ull, 5ull);
    SWV("heatbath  rolled nu loop (round 1)", 1, (lq_sweep4_rolled_kernel<128, 3, 0>), g, W, mu, par, 0, 0, 6.0, 0x777ull, 5ull);
    SWV("overrelax product kernel (pipelined staples)", 2, (lq_sweep4_kernel<128, 3, 1>), g, W, mu, par, 0, 0, 6.0, 0x777ull, 5ull);
    SWV("overrelax rolled nu loop (round 1)", 3, (lq_sweep4_rolled_kernel<128, 3, 1>), g, W, mu, par, 0, 0, 6.0, 0x777ull, 5ull);
    SWV("overrelax su2-subgroups product kernel", 4, (lq_sweep4_kernel<128, 3, 1>), g, W, mu, par, 0, 2, 6.0, 0x777ull, 5ull);
    SWV("overrelax su2-subgroups rolled", 5, (lq_sweep4_rolled_kernel<128, 3, 1>), g, W, mu, par, 0, 2, 6.0, 0x777ull, 5ull);
    SWV("staple phase alone (pipelined, 168 regs)", -1, (lq_sweep4_staples_only_kernel<128, 3, 1>), g, W, U2, mu, par);
    SWV("staple phase alone (pipelined, 255 regs)", -1, (lq_sweep4_staples_only_kernel<128, 2, 1>), g, W, U2, mu, par);
    SWV("staple phase alone (pipelined, 128 regs)", -1, (lq_sweep4_staples_only_kernel<128, 4, 1>), g, W, U2, mu, par);
    // warp-specialised sub-step (staple warps + rule warps, persistent blocks): warp and register splits
    int nsm = 148;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    const int ntasks = (int)((g.vol / 2 + 31) / 32);
    const unsigned gws = (unsigned)(ntasks < nsm ? ntasks : nsm);
#define SWS(NAME, KIND, NPW, NCW, PREG, CREG, ORK)                                                                   \
  CK(cudaFuncSetAttribute(lq_sweep4ws_kernel<KIND, NPW, NCW, PREG, CREG>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                          LqWsCfg<NPW, NCW>::SMEM));                                                                   \
  sv.push_back({NAME, [&](int mu, int par) {                                                                         \
                  lq_sweep4ws_kernel<KIND, NPW, NCW, PREG, CREG><<<gws, LqWsCfg<NPW, NCW>::THREADS, LqWsCfg<NPW, NCW>::SMEM>>>( \
                      g, W, mu, par, 0, ORK, 6.0, 0x777ull, 5ull, ntasks);                                           \
                  CK(cudaGetLastError());                                                                            \
                },                                                                                                   \
                (const void*)lq_sweep4ws_kernel<KIND, NPW, NCW, PREG, CREG>, LqWsCfg<NPW, NCW>::THREADS,               \
                KIND == 0 ? 0 : (ORK == 2 ? 4 : 2), LqWsCfg<NPW, NCW>::SMEM})
    // register splits must balance: NPW (PREG - R0) <= NCW (R0 - CREG), R0 = the launch allocation (128 at 512 threads,
    // 96 at 640): setmaxnreg.inc only draws from what the block's own warps released
    SWS("heatbath  warp-spec. 8+8 warps 168/88", 0, 8, 8, 168, 88, 0);
    SWS("heatbath  warp-spec. 8+8 warps 152/104", 0, 8, 8, 152, 104, 0);
    SWS("heatbath  warp-spec. 12+8 warps 112/72", 0, 12, 8, 112, 72, 0);
    SWS("heatbath  warp-spec. 12+8 warps 104/80", 0, 12, 8, 104, 80, 0);
    SWS("heatbath  warp-spec. 12+4 warps 136/104", 0, 12, 4, 136, 104, 0);
    SWS("heatbath  warp-spec. 12+4 warps 128/128", 0, 12, 4, 0, 0, 0);
    SWS("overrelax warp-spec. 8+8 warps 128/128", 1, 8, 8, 0, 0, 0);
    SWS("overrelax su2-subgroups warp-spec. 12+4 144/80", 1, 12, 4, 144, 80, 2);
    SWS("overrelax su2-subgroups warp-spec. 8+8 168/88", 1, 8, 8, 168, 88, 2);
    printf("%-52s %5s %6s %4s %9s %9s %9s\n", "sweep variant (8 sub-steps)", "regs", "local", "w/SM", "ms/sweep", "GB/s(alg)",
           "max|dU|");
    std::vector<double> hW(ub / 8), hR[3];
    for (auto& v : sv) {
      if (filter[0] && v.name.find(filter) == std::string::npos && v.name.find("product kernel") == std::string::npos) continue;
      cudaFuncAttributes at;
      CK(cudaFuncGetAttributes(&at, v.func));
      int occ = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.func, v.block, v.smem));
      CK(cudaMemcpy(W, U, ub, cudaMemcpyDeviceToDevice));
      for (int mu = 0; mu < 4; ++mu)
        for (int par = 0; par < 2; ++par) v.run(mu, par);
      CK(cudaDeviceSynchronize());
      // one sweep from the common start: the first variant of a rule family is the reference of the ones after it
      double dmax = -1.0;
      if (v.family >= 0) {
        CK(cudaMemcpy(hW.data(), W, ub, cudaMemcpyDeviceToHost));
        const int f = v.family / 2;
        if (v.family % 2 == 0 && hR[f].empty()) {
          hR[f] = hW;
        } else if (!hR[f].empty()) {
          dmax = 0.0;
          for (size_t i = 0; i < hW.size(); ++i) dmax = fmax(dmax, fabs(hW[i] - hR[f][i]));
        }
      }
      CK(cudaEventRecord(e0));
      for (int r = 0; r < reps; ++r)
        for (int mu = 0; mu < 4; ++mu)
          for (int par = 0; par < 2; ++par) v.run(mu, par);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ms /= reps;
      printf("%-52s %5d %6zu %4d %9.4f %9.1f %9.2e\n", v.name.c_str(), at.numRegs, at.localSizeBytes, occ * v.block / 32, ms,
             1296.0 * nl / ms / 1e6, dmax);
      fflush(stdout);
    }
    CK(cudaFree(W));
  }
  return 0;
}
