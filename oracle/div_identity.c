// TEST INFRASTRUCTURE (oracle/): exhaustive CPU check of the identity the fused activation-quant kernel relies on
// (qqq_b200/csrc/act_quant.cu, restating QQQ/gptq/qlinear/qlinear_marlin.py:265-268 `x / quant_scale`):
// for every finite fp16 x and every positive finite fp16-valued s,
//   q1 = fma(fma(-s, q0, x), r, q0)  with r = RN(1/s), q0 = RN(x*r)   equals   RN(x / s)   bit for bit (fp32),
// except for the sign of a zero quotient (x = -0), which the int8 conversion erases.
// Build: gcc -O2 -mfma -ffp-contract=off -fopenmp div_identity.c -lm   (tests/test_act_quant_identity.py does it)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static float h2f(uint16_t h) {
  uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023, u;
  if (e == 0) {
    if (m == 0) u = s << 31;
    else { int sh = 0; while (!(m & 1024)) { m <<= 1; ++sh; } m &= 1023; u = (s << 31) | ((127 - 15 - sh + 1) << 23) | (m << 13); }
  } else if (e == 31) u = (s << 31) | 0x7F800000u | (m << 13);
  else u = (s << 31) | ((e - 15 + 127) << 23) | (m << 13);
  float f; memcpy(&f, &u, 4); return f;
}
int main(void) {
  long long bad = 0, bad_q8 = 0, total = 0;
  #pragma omp parallel for schedule(dynamic, 64) reduction(+:bad,bad_q8,total)
  for (int sb = 1; sb < 0x7C00; ++sb) {
    const float s = h2f((uint16_t)sb);
    const float r = 1.0f / s;
    for (int xb = 0; xb < 0x10000; ++xb) {
      if (((xb >> 10) & 31) == 31) continue;
      const float x = h2f((uint16_t)xb);
      const float q0 = x * r;
      const float e = fmaf(-s, q0, x);
      const float q1 = fmaf(e, r, q0);
      const float q = x / s;
      uint32_t a, b; memcpy(&a, &q1, 4); memcpy(&b, &q, 4);
      ++total;
      if (a != b && xb != 0x8000) {
        ++bad;
        float ra = rintf(q1), rb = rintf(q);
        ra = fminf(fmaxf(ra, -128.f), 127.f); rb = fminf(fmaxf(rb, -128.f), 127.f);
        if (ra != rb) ++bad_q8;
      }
    }
  }
  printf("pairs=%lld  fp32 quotient mismatches=%lld  int8 mismatches=%lld\n", total, bad, bad_q8);
  return bad_q8 != 0;
}
