// fft_registry.h -- host-side table of the compiled FFT kernel instantiations.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "fft_kernel_v2.cuh"

namespace d2d {

constexpr int kMaxDevices = 64; // per-device launch caches

enum FftKind { KIND_LINE = 0 /* TX = 1, contiguous lines */, KIND_TILE = 1 /* TX > 1, strided lines */, KIND_TILE_WIDE = 2 /* 2x wider tile rows */ };

struct FftKernelInfo {
   int n, f64, kind, mode, pairvec, line_in; // line_in: shared memory laid out for inputs contiguous along the transform axis
   int tx, ly, threads, minb;
   size_t smem;
   int tw_total;                 // complex twiddle entries expected in FftArgs::tw
   int npass, radix[4];
   const void *func;             // for cudaFuncSetAttribute / occupancy queries
   cudaError_t (*launch)(const FftArgs &, cudaStream_t);
   // v2 (TMA-staged) kernels: fft_kernel_v2.cuh
   int v2 = 0, inl = 0;          // inl: IN_TILE / IN_LINE
   int rows = 0, rows_early = 0, row_bytes = 0; // IN_TILE landing geometry (Geom2)
   int merged = 0;               // IN_TILE: the LY sub-tiles land together (one box row = ly * row_bytes)
   cudaError_t (*launch2)(const FftArgs2 &, const TmapPack &, cudaStream_t) = nullptr;
};

void fft_register(const FftKernelInfo &);
// exact lookup; nullptr if this (n, dtype, kind, mode, pairvec) was not compiled
const FftKernelInfo *fft_find(int n, int f64, int kind, int mode, int pairvec, int line_in = 0);
const FftKernelInfo *fft_find_v2(int n, int f64, int mode, int inl, int row_bytes = 64, int merged = 0);
int fft_registry_size();
const FftKernelInfo *fft_registry_at(int i);

} // namespace d2d
