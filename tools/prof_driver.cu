// Torch-free driver for ncu captures of the conv kernels: links libcfun_b200.so, feeds it cudaMalloc'ed buffers through the
// C ABI (exactly what the Python binding does), runs forward / data gradient / weight gradient of one conv shape.
//   nvcc -O2 -o tools/prof_driver.bin tools/prof_driver.cu -Lcfun_b200/lib -lcfun_b200 -Xlinker -rpath='$ORIGIN/../cfun_b200/lib'
//   tools/prof_driver.bin unet|rpn|l3|s2 [reps]
// (ncu attaches in seconds here; under `python` it spends minutes patching torch's modules before the first launch.)
#include "../include/cfun_b200.h"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
  const char* which = argc > 1 ? argv[1] : "unet";
  const int reps = argc > 2 ? atoi(argv[2]) : 2;
  cfun_conv3d_desc d;
  memset(&d, 0, sizeof(d));
  int N = 4, Ci = 40, S = 96, Co = 40, st = 1;
  if (!strcmp(which, "rpn")) { N = 1; Ci = 128; S = 32; Co = 256; }
  else if (!strcmp(which, "l3")) { N = 4; Ci = 80; S = 48; Co = 80; }
  else if (!strcmp(which, "thin")) { N = 4; Ci = 20; S = 96; Co = 20; }
  else if (!strcmp(which, "s2")) { N = 4; Ci = 20; S = 96; Co = 40; st = 2; }
  // small shapes for compute-sanitizer runs (same kernels, seconds instead of minutes)
  else if (!strcmp(which, "small")) { N = 2; Ci = 24; S = 24; Co = 40; }
  else if (!strcmp(which, "smallhx")) { N = 1; Ci = 96; S = 16; Co = 160; }
  else if (!strcmp(which, "smalls2")) { N = 1; Ci = 20; S = 32; Co = 40; st = 2; }
  else if (!strcmp(which, "smallwide")) { N = 1; Ci = 176; S = 16; Co = 48; }
  d.N = N; d.Cin = Ci; d.Din = d.Hin = d.Win = S; d.Cout = Co;
  d.kD = d.kH = d.kW = 3; d.sD = d.sH = d.sW = st; d.pD = d.pH = d.pW = 1;
  d.Dout = d.Hout = d.Wout = (S + 2 - 3) / st + 1;
  const size_t nx = (size_t)N * S * S * S * Ci, ny = (size_t)N * d.Dout * d.Hout * d.Wout * Co, nw = (size_t)Co * Ci * 27;
  std::vector<float> hx(nx), hy(ny), hw(nw);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
  for (auto& v : hx) v = rnd();
  for (auto& v : hy) v = rnd();
  for (auto& v : hw) v = 0.05f * rnd();
  float *x, *y, *w, *dx, *dw;
  CK(cudaMalloc(&x, nx * 4)); CK(cudaMalloc(&dx, nx * 4)); CK(cudaMalloc(&y, ny * 4)); CK(cudaMalloc(&w, nw * 4)); CK(cudaMalloc(&dw, nw * 4));
  CK(cudaMemcpy(x, hx.data(), nx * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w, hw.data(), nw * 4, cudaMemcpyHostToDevice));
  size_t wsb = 0;
  for (int pass = 0; pass < 3; ++pass) { size_t b = cfun_conv3d_workspace_size(&d, pass, CFUN_CONV_ALGO_AUTO); if (b > wsb) wsb = b; }
  void* ws;
  CK(cudaMalloc(&ws, wsb + 4096));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[3] = {"fwd", "dgrad", "wgrad"};
  for (int r = 0; r < reps; ++r) {
    for (int pass = 0; pass < 3; ++pass) {
      int rc;
      cudaEventRecord(e0);
      if (pass == 0) rc = cfun_conv3d_fwd(&d, x, w, nullptr, y, 0, CFUN_CONV_ALGO_AUTO, ws, wsb + 4096, nullptr);
      else if (pass == 1) rc = cfun_conv3d_bwd_data(&d, y, w, dx, CFUN_CONV_ALGO_AUTO, ws, wsb + 4096, nullptr);
      else rc = cfun_conv3d_bwd_weight(&d, x, y, dw, nullptr, CFUN_CONV_ALGO_AUTO, ws, wsb + 4096, nullptr);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("%s %s N%d %d->%d @%d s%d: rc %d (%s) %.3f ms algo %d\n", which, names[pass], N, Ci, Co, S, st, rc, rc ? cfun_last_error() : "ok", ms,
             cfun_conv3d_pick_algo(&d, pass));
      if (pass == 0 && r == 0) CK(cudaMemcpy(y, hy.data(), ny * 4, cudaMemcpyHostToDevice));   // dY for the backward passes
    }
  }
  return 0;
}
