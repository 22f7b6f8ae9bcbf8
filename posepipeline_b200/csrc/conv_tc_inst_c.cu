// conv_tc kernel instantiations, part c (see conv_tc_kernel.cuh / conv_tc.cu)
#include "conv_tc_kernel.cuh"

// (MT, NC, TAPS, KC, CG, SETS) instantiations: MT*NC in {16..128} columns; 3x3 (9 taps), 2x2 (4 taps) and 1-D (3 taps) stages hold
// one 16-channel chunk, 1x1 stages hold up to four; CTA-pair forms (CG = 2) for N >= 32; two epilogue sets for the 3x3 and 1-D kinds.
// The instantiations are spread over conv_tc_inst_{a..f}.cu (one layer kind each) so that they compile in parallel.
#define TCK(mt, nc, taps, kc, cg, sets) if (MT == mt && NC == nc && TAPS == taps && KC == kc && CG == cg && SETS == sets) return conv_tc_kernel<(mt) * (nc) / 16, mt, taps, kc, cg, sets>;
#define TCK_SHAPES(taps, kc, sets) \
  TCK(1, 16, taps, kc, 1, sets) \
  TCK(1, 32, taps, kc, 1, sets) TCK(2, 32, taps, kc, 1, sets) TCK(1, 48, taps, kc, 1, sets) TCK(2, 48, taps, kc, 1, sets) TCK(1, 64, taps, kc, 1, sets) \
  TCK(2, 64, taps, kc, 1, sets) TCK(1, 80, taps, kc, 1, sets) TCK(1, 96, taps, kc, 1, sets) TCK(1, 128, taps, kc, 1, sets) \
  TCK(1, 32, taps, kc, 2, sets) TCK(2, 32, taps, kc, 2, sets) TCK(1, 48, taps, kc, 2, sets) TCK(2, 48, taps, kc, 2, sets) TCK(1, 64, taps, kc, 2, sets) \
  TCK(2, 64, taps, kc, 2, sets) TCK(1, 80, taps, kc, 2, sets) TCK(1, 96, taps, kc, 2, sets) TCK(1, 128, taps, kc, 2, sets)

TcKernelFn tc_kernel_for_c(int MT, int NC, int TAPS, int KC, int CG, int SETS) {
  TCK_SHAPES(4, 1, 1) TCK_SHAPES(4, 1, 3) TCK_SHAPES(4, 1, 4)
  return nullptr;
}
