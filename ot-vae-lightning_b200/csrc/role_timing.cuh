// Per-role clock accounting for the warp-specialised kernels (-DOTK_SH_TIMING): where each warp role spends its cycles,
// printed by lane 0 of one warp per role of CTAs 0 and 77 at the end of the kernel; compiled out otherwise.
#pragma once
#ifdef OTK_SH_TIMING
#include <cstdio>
#define S2_T0 long long s2_prev = clock64(), s2_a = 0, s2_b = 0, s2_c = 0, s2_d = 0, s2_e = 0; const long long s2_start = s2_prev; int s2_n = 0;
#define S2_TICK(acc) { const long long s2_now = clock64(); acc += s2_now - s2_prev; s2_prev = s2_now; }
#define S2_COUNT ++s2_n;
#define S2_REPORT(role, na, nb, nc, nd, ne) \
  if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77) && s2_n > 0) \
    printf("cta %3d %-10s steps %4d: total %6lld | " na " %5lld | " nb " %5lld | " nc " %5lld | " nd " %5lld | " ne " %5lld (clk per step)\n", \
           (int)blockIdx.x, role, s2_n, (clock64() - s2_start) / s2_n, s2_a / s2_n, s2_b / s2_n, s2_c / s2_n, s2_d / s2_n, s2_e / s2_n);
#else
#define S2_T0
#define S2_TICK(acc)
#define S2_COUNT
#define S2_REPORT(role, na, nb, nc, nd, ne)
#endif
