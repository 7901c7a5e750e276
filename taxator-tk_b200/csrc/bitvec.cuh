// Multi-word integer helpers of the bit-vector edit-distance kernel (myers3.cuh): carry-chained adds
// (add.cc / addc.cc -> IADD3.X) over the W words of a lane, and explicit 3-input LOP3s.
#pragma once
#include "common.cuh"

namespace trpa {

// sum[i] = a[i] + b[i] + carry chain; carry-in = (cin + K) overflow, carry-out returned as 0/1.
// K = 0x80000000 when cin carries the flag in bit 31, K = 0xffffffff when cin is 0/1.
template <int N>
struct AddChain;

template <>
struct AddChain<1> {
  static __device__ __forceinline__ u32 run(u32* s, const u32* a, const u32* b, u32 cin, u32 K) {
    u32 co;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %2, %3;\n\t"
        "addc.cc.u32 %0, %4, %5;\n\t"
        "addc.u32 %1, 0, 0;\n\t}"
        : "=r"(s[0]), "=r"(co)
        : "r"(cin), "r"(K), "r"(a[0]), "r"(b[0]));
    return co;
  }
};
template <>
struct AddChain<2> {
  static __device__ __forceinline__ u32 run(u32* s, const u32* a, const u32* b, u32 cin, u32 K) {
    u32 co;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %3, %4;\n\t"
        "addc.cc.u32 %0, %5, %6;\n\t"
        "addc.cc.u32 %1, %7, %8;\n\t"
        "addc.u32 %2, 0, 0;\n\t}"
        : "=r"(s[0]), "=r"(s[1]), "=r"(co)
        : "r"(cin), "r"(K), "r"(a[0]), "r"(b[0]), "r"(a[1]), "r"(b[1]));
    return co;
  }
};
template <>
struct AddChain<3> {
  static __device__ __forceinline__ u32 run(u32* s, const u32* a, const u32* b, u32 cin, u32 K) {
    u32 co;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %4, %5;\n\t"
        "addc.cc.u32 %0, %6, %7;\n\t"
        "addc.cc.u32 %1, %8, %9;\n\t"
        "addc.cc.u32 %2, %10, %11;\n\t"
        "addc.u32 %3, 0, 0;\n\t}"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(co)
        : "r"(cin), "r"(K), "r"(a[0]), "r"(b[0]), "r"(a[1]), "r"(b[1]), "r"(a[2]), "r"(b[2]));
    return co;
  }
};
template <>
struct AddChain<4> {
  static __device__ __forceinline__ u32 run(u32* s, const u32* a, const u32* b, u32 cin, u32 K) {
    u32 co;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %5, %6;\n\t"
        "addc.cc.u32 %0, %7, %8;\n\t"
        "addc.cc.u32 %1, %9, %10;\n\t"
        "addc.cc.u32 %2, %11, %12;\n\t"
        "addc.cc.u32 %3, %13, %14;\n\t"
        "addc.u32 %4, 0, 0;\n\t}"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(co)
        : "r"(cin), "r"(K), "r"(a[0]), "r"(b[0]), "r"(a[1]), "r"(b[1]), "r"(a[2]), "r"(b[2]),
          "r"(a[3]), "r"(b[3]));
    return co;
  }
};

// chain over W words in chunks of <= 4
template <int W>
__device__ __forceinline__ u32 add_words(u32 (&s)[W], const u32 (&a)[W], const u32 (&b)[W], u32 cin_msb) {
  u32 c = cin_msb;
  u32 K = 0x80000000u;
#pragma unroll
  for (int w0 = 0; w0 < W; w0 += 4) {
    if (W - w0 >= 4) c = AddChain<4>::run(&s[w0], &a[w0], &b[w0], c, K);
    else if (W - w0 == 3) c = AddChain<3>::run(&s[w0], &a[w0], &b[w0], c, K);
    else if (W - w0 == 2) c = AddChain<2>::run(&s[w0], &a[w0], &b[w0], c, K);
    else c = AddChain<1>::run(&s[w0], &a[w0], &b[w0], c, K);
    K = 0xffffffffu;
  }
  return c;  // 0/1
}

template <int LUT>
__device__ __forceinline__ u32 lop3(u32 a, u32 b, u32 c) {
  u32 d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return d;
}

}  // namespace trpa
