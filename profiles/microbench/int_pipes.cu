// int_pipes.cu — per-SMSP reciprocal throughput of the integer ops the placement kernel is made of.
// One CTA per SM, W warps, each lane runs 8 independent dependency chains of one op.
// Prints cycles per warp-instruction per SM sub-partition (SMSP) at 4, 8 and 16 warps/SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

template <int OP>
__global__ void k(int* out, long long* cyc, int a0, int b0) {
    int r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a0 + j + threadIdx.x;
    int b = b0, c = b0 * 3 + 1;
    int q[8], q2[8], bb[8]; unsigned x = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { q[j] = j; q2[j] = 2 * j; bb[j] = 0x7fff7fff; }
    double d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = (double)(j + threadIdx.x);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) r[j] = __dp4a(r[j], b, r[j]);                       // IDP.4A
            if (OP == 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(c));   // IMAD
            if (OP == 2) asm volatile("add.s32 %0, %0, %1;" : "+r"(r[j]) : "r"(b));                  // IADD3
            if (OP == 3) asm volatile("min.s32 %0, %0, %1;" : "+r"(r[j]) : "r"(b + i));              // VIMNMX
            if (OP == 4) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; selp.s32 %0, %1, %0, p;}" : "+r"(r[j]) : "r"(b + i));  // ISETP+SEL
            if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(c));     // PRMT
            if (OP == 6) asm volatile("{.reg .pred p; setp.eq.s32 p, %1, %2; @p add.f64 %0, %0, %3;}" : "+d"(d[j]) : "r"(r[j]), "r"(b), "d"(1.5));  // ISETP + @DADD
            if (OP == 7) asm volatile("add.f64 %0, %0, %1;" : "+d"(d[j]) : "d"(1.5));                // DADD
            if (OP == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[j]) : "r"(b), "r"(c));  // LOP3
            if (OP == 9) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(c)); // SHF
            if (OP == 10) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; @p add.s32 %0, %0, %2;}" : "+r"(r[j]) : "r"(b + i), "r"(c)); // ISETP+@IADD
            if (OP == 11) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(r[j]) : "r"(b));               // packed add
            if (OP == 12) asm volatile("min.s16x2 %0, %0, %1;" : "+r"(r[j]) : "r"(b + i));           // packed min
            if (OP == 13) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(r[j]) : "r"(b), "r"(c));  // IMAD as add (a*b const + r)
            if (OP == 14) asm volatile("{.reg .pred p; setp.eq.s32 p, %0, %1; @p add.s32 %0, %0, %2;  min.s32 %0, %0, %1;}" : "+r"(r[j]) : "r"(b + i), "r"(c)); // setp+@add+min
            if (OP == 15) r[j] += __popc(r[j] ^ b);                                                    // POPC (+LOP,+IADD)
            if (OP == 16) {  // packed pass-1 step for one pair of reads: PRMT + VIADD.16x2 + VIMNMX.S16x2(+2 preds) + 2 @IADD + LOP3
                unsigned dlt, nb;
                asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(dlt) : "r"(b), "r"(c), "r"(i));
                asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r[j]) : "r"(dlt));
                asm volatile("{.reg .pred ph, pl; .reg .u16 a0, a1, m0, m1;\n\t"
                             "min.s16x2 %0, %3, %4;\n\t"
                             "mov.b32 {m0, m1}, %0; mov.b32 {a0, a1}, %3;\n\t"
                             "setp.eq.s16 pl, m0, a0; setp.eq.s16 ph, m1, a1;\n\t"
                             "@pl add.s32 %1, %1, %5;\n\t"
                             "@ph add.s32 %2, %2, %5;}"
                             : "=r"(nb), "+r"(q[j]), "+r"(q2[j]) : "r"(r[j]), "r"(bb[j]), "r"(c));
                x |= nb ^ bb[j];
                bb[j] = nb;
            }
            if (OP == 18) r[j] = __viaddmin_s16x2(r[j], b, c + i);                                      // VIADDMNMX.S16x2
            if (OP == 19) r[j] = __umulhi(r[j], 0x10000u) + b;                                           // IMAD.HI (shift by 16 on the FMA pipe)
            if (OP == 20) asm volatile("{.reg .pred p; setp.eq.s32 p, %1, 0x7fffffff; @p prmt.b32 %0, %0, %1, %2; @p prmt.b32 %0, %0, %2, %1; @p prmt.b32 %0, %1, %0, %2; @p prmt.b32 %0, %2, %0, %1;}" : "+r"(r[j]) : "r"(b), "r"(c));  // 4 predicated-off PRMT
            if (OP == 21) asm volatile("{.reg .pred p; setp.ne.s32 p, %1, 3; selp.b32 %0, %0, %2, p;}" : "+r"(r[j]) : "r"(b), "r"(c));   // SEL
            if (OP == 22) asm volatile("prmt.b32 %0, %0, %1, %2; min.s16x2 %0, %0, %1;" : "+r"(r[j]) : "r"(b), "r"(c));   // PRMT + VIMNMX pair
            if (OP == 23) asm volatile("prmt.b32 %0, %0, %1, %2; mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(c));   // PRMT + IMAD pair
            if (OP == 17) asm volatile("{.reg .pred ph, pl; .reg .u16 a0, a1, m0, m1; .reg .b32 t;\n\t"
                             "min.s16x2 t, %0, %2;\n\t"
                             "mov.b32 {m0, m1}, t; mov.b32 {a0, a1}, %0;\n\t"
                             "setp.eq.s16 pl, m0, a0; setp.eq.s16 ph, m1, a1;\n\t"
                             "@pl add.s32 %1, %1, %3;\n\t"
                             "@ph add.s32 %0, %0, %3;}"
                             : "+r"(r[j]), "+r"(q[j]) : "r"(b + i), "r"(c));   // VIMNMX.S16x2 w/ preds + 2 @IADD
        }
    }
    long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += r[j] + (int)d[j] + q[j] + q2[j] + bb[j];
    acc += x;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int n_instr_per_iter) {
    int* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(int));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    printf("%-26s", name);
    for (int warps : {4, 8, 16, 32}) {
        k<OP><<<148, warps * 32>>>(out, cyc, 1, 3);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        // warp-instructions per SMSP = warps/4 * ITER * 8 * n_instr
        double per = avg / ((warps / 4.0) * ITER * 8.0 * n_instr_per_iter);
        printf("  w%-2d %6.2f", warps, per);
    }
    printf("   cyc/warp-instr/SMSP (%d instr per op)\n", n_instr_per_iter);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("IDP.4A", 1);
    run<1>("IMAD", 1);
    run<13>("IMAD (r += b*c)", 1);
    run<2>("IADD", 1);
    run<3>("min.s32 (VIMNMX)", 1);
    run<4>("ISETP+SEL", 2);
    run<10>("ISETP+@IADD", 2);
    run<14>("ISETP+@IADD+MIN", 3);
    run<5>("PRMT", 1);
    run<8>("LOP3", 1);
    run<9>("SHF", 1);
    run<6>("ISETP+@DADD", 2);
    run<7>("DADD", 1);
    run<11>("add.u16x2", 1);
    run<12>("min.s16x2", 1);
    run<15>("LOP+POPC+IADD", 3);
    run<18>("VIADDMNMX.S16x2", 1);
    run<19>("IMAD.HI + IADD", 2);
    run<20>("ISETP + 4 pred-off PRMT", 5);
    run<21>("ISETP + SEL", 2);
    run<22>("PRMT + VIMNMX", 2);
    run<23>("PRMT + IMAD", 2);
    run<17>("VIMNMX.S16x2+P,2x@IADD", 3);
    run<16>("packed pass-1 pair step", 6);
    return 0;
}
