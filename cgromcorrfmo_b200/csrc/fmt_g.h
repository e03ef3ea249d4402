// "%g" (6 significant digits) of a float, byte-identical to printf/std::ostream (the reference prints every .time and
// .mtx value with `stream << float`, reference src/FMOparse.cpp:162,216), several times faster than snprintf.
//
// Fast path: scale |v| by an exactly representable power of ten so that the 6 digits are the integer part, round to
// nearest, and print.  The scaling multiplication rounds once (relative 1.1e-16), so the result can only differ from the
// exactly-rounded decimal expansion when the scaled value sits within ~1e-10 of a rounding tie; those cases (about one
// in 10^7), non-finite values and exponents outside the exact-power range go to snprintf.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>

namespace cgc {

inline int fmt_g6(char* p, float f) {
  static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                    1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const double v = (double)f;
  if (f == 0.0f) {
    if (std::signbit(f)) { p[0] = '-'; p[1] = '0'; return 2; }
    p[0] = '0';
    return 1;
  }
  const double a = std::fabs(v);
  if (!(a >= 1e-16 && a < 1e16)) return snprintf(p, 32, "%g", v);  // also catches NaN/Inf
  // decimal exponent: estimate from the binary exponent, then correct by at most one (thresholds are only used to order)
  static const double kThr[37] = {1e-18, 1e-17, 1e-16, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6,
                                  1e-5,  1e-4,  1e-3,  1e-2,  1e-1,  1e0,   1e1,   1e2,   1e3,  1e4,  1e5,  1e6,
                                  1e7,   1e8,   1e9,   1e10,  1e11,  1e12,  1e13,  1e14,  1e15, 1e16, 1e17, 1e18};
  unsigned long long bits;
  memcpy(&bits, &a, 8);
  const int e2 = (int)((bits >> 52) & 0x7ff) - 1023;         // a = 1.m * 2^e2
  int e10 = (e2 * 78913) >> 18;                               // floor(e2 * log10(2)) for |e2| < 2^12
  if (a < kThr[e10 + 18]) e10--;
  else if (a >= kThr[e10 + 19]) e10++;
  const int k = 5 - e10;                    // scaled = a * 10^k has 6 integer digits
  const double scaled = k >= 0 ? a * kPow10[k] : a / kPow10[-k];
  if (!(scaled >= 99999.0 && scaled < 1000001.0)) return snprintf(p, 32, "%g", v);  // exponent estimate off by one ulp
  const double fl = (double)(long)scaled;
  const double frac = scaled - fl;
  if (std::fabs(frac - 0.5) < 1e-6 || fl < 100000.0) return snprintf(p, 32, "%g", v);  // too close to a tie / boundary
  long r = (long)fl + (frac > 0.5 ? 1 : 0);
  if (r >= 1000000) { r = 100000; e10++; }
  if (r < 100000) return snprintf(p, 32, "%g", v);
  char d[6];
  for (int i = 5; i >= 0; i--) { d[i] = (char)('0' + r % 10); r /= 10; }
  int nd = 6;
  while (nd > 1 && d[nd - 1] == '0') nd--;
  char* q = p;
  if (v < 0) *q++ = '-';
  if (e10 < -4 || e10 >= 6) {
    *q++ = d[0];
    if (nd > 1) { *q++ = '.'; memcpy(q, d + 1, (size_t)(nd - 1)); q += nd - 1; }
    *q++ = 'e';
    int e = e10;
    if (e < 0) { *q++ = '-'; e = -e; } else { *q++ = '+'; }
    if (e >= 100) { *q++ = (char)('0' + e / 100); e %= 100; }
    *q++ = (char)('0' + e / 10);
    *q++ = (char)('0' + e % 10);
  } else if (e10 >= 0) {
    const int ni = e10 + 1;                 // digits before the point
    if (nd <= ni) {
      memcpy(q, d, (size_t)nd); q += nd;
      for (int i = nd; i < ni; i++) *q++ = '0';
    } else {
      memcpy(q, d, (size_t)ni); q += ni;
      *q++ = '.';
      memcpy(q, d + ni, (size_t)(nd - ni)); q += nd - ni;
    }
  } else {
    *q++ = '0'; *q++ = '.';
    for (int i = 0; i < -e10 - 1; i++) *q++ = '0';
    memcpy(q, d, (size_t)nd); q += nd;
  }
  return (int)(q - p);
}

}  // namespace cgc
