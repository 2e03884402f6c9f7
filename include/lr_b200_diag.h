/* lr_b200_diag.h — measurement hooks and tcgen05 micro-benchmarks.  NOT part of the product library: these symbols
 * exist only in liblr_b200_diag.so (the same sources compiled with -DLR_DIAG plus csrc/diag/), which tools/ load.
 * Unlike lr_b200.h they keep process-global state (that is what a hook is). */
#ifndef LR_B200_DIAG_H_
#define LR_B200_DIAG_H_
#include "lr_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* -------- diagnostics ------------------------------------------------------------------------ */
/* SM cycles for `iters` back-to-back tcgen05.mma of one shape with operands resident in shared memory
 * (tools/umma_table.py): the measured per-instruction cost that tile-orientation choices are based on.
 * Synchronous (the only entry point that is).                                                    */
/* Device buffer (148*8 int64, or NULL to stop) that subsequent lr_conv3d_fwd launches fill with the
 * cycles each warp role spent waiting on its barriers (tools/conv_waits.py).                      */
void lr_conv3d_set_debug(long long* device_buffer);
/* Diagnostics only (results become wrong): bit 0 skips the epilogue work, bit 1 loads each weight stage once,
 * bit 2 loads each input chunk set once — shows which role limits a layer (tools/conv_waits.py --skip). */
void lr_conv3d_set_debug_skip(int mask);
long long lr_umma_microbench(int M, int N, int row_bytes_a, int row_bytes_b, int a_major, int b_major,
                             int n_acc, int a_tiles, int iters, int a_shift_rows, void* stream);

/* Same, for a trip of 8 MMAs (M = 128, K-major 64-byte rows) with individual N, accumulator column and B row-group
 * offsets: what differently shaped MMAs on overlapping accumulator ranges cost (tools/umma_pattern.py);
 * commit_every > 0 adds a tcgen05.commit after every that many MMAs (multiple of 8).             */
long long lr_umma_pattern_bench(const int* n, const int* dcol, const int* bblk, int iters, int commit_every,
                                void* stream);

/* Cycles for tiles*n_ops*KS MMAs (M = 128, K-major 64-byte rows) issued by ONE thread of a `threads`-thread block
 * from a loop whose descriptors advance by loop-carried adds (variant 0) or stay constant (variant 1). */
long long lr_umma_issue_bench(int N, int KS, int tiles, int n_ops, int a_step_bytes, int d_step, int threads,
                              int variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LR_B200_DIAG_H_ */
