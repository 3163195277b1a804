// Microbenchmark: issue throughput of scalar FMUL/FADD vs packed FFMA2/FADD2/FMUL2 on sm_100a.
// Every product depends on a loop-carried value so nothing is hoisted.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o mb_f32x2 mb_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;}
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
constexpr int N = 8;
// mode 0: FMUL+FADD   1: FMUL only   2: FADD only   3: FFMA only
template <int MODE> __global__ void k_scalar(float* out, float f, int iters){
  float a[N];
  for(int t=0;t<N;++t) a[t]=threadIdx.x*1e-3f+t;
  for(int i=0;i<iters;++i){
#pragma unroll
    for(int t=0;t<N;++t){
      if (MODE==0) a[t]=__fadd_rn(a[t], __fmul_rn(a[(t+1)%N], f));
      if (MODE==1) a[t]=__fmul_rn(a[(t+1)%N], f);
      if (MODE==2) a[t]=__fadd_rn(a[t], a[(t+1)%N]);
      if (MODE==3) a[t]=__fmaf_rn(a[(t+1)%N], f, a[t]);
    }
  }
  float s=0; for(int t=0;t<N;++t) s+=a[t];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// mode 0: FFMA2(-0)+FADD2  1: FFMA2 only  2: FADD2 only  3: FMUL2 + 2 FADD  4: FMUL2 only
template <int MODE> __global__ void k_packed(float* out, float f, int iters, u64 nz){
  u64 a[N];
  for(int t=0;t<N;++t) a[t]=pk(threadIdx.x*1e-3f+t, threadIdx.x*2e-3f+t);
  u64 ff = pk(f,f);
  for(int i=0;i<iters;++i){
#pragma unroll
    for(int t=0;t<N;++t){
      if (MODE==0) a[t]=add2(a[t], fma2(a[(t+1)%N], ff, nz));
      if (MODE==1) a[t]=fma2(a[(t+1)%N], ff, a[t]);
      if (MODE==2) a[t]=add2(a[t], a[(t+1)%N]);
      if (MODE==3){ u64 p = mul2(a[(t+1)%N], ff); float p0,p1,c0,c1; upk(p,p0,p1); upk(a[t],c0,c1); a[t]=pk(__fadd_rn(c0,p0), __fadd_rn(c1,p1)); }
      if (MODE==4) a[t]=mul2(a[(t+1)%N], ff);
    }
  }
  float s=0; for(int t=0;t<N;++t){ float x,y; upk(a[t],x,y); s+=x+y; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template <typename F> float timeit(F f){ cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); f(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); return ms; }
int main(){
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, sizeof(float)*sms*8*256);
  int iters=20000; int blocks=sms*8, threads=256;
  u64 nz = 0x8000000080000000ull;
  double lanes = (double)blocks*threads*iters*N;  // scalar element-updates per launch
  const char* sn[4]={"FMUL+FADD","FMUL","FADD","FFMA"};
  float ms;
  ms=timeit([&]{k_scalar<0><<<blocks,threads>>>(out,1.0001f,iters);}); printf("scalar %-12s %.3f ms  %.1f G elem-updates/s\n", sn[0], ms, lanes/(ms*1e-3)/1e9);
  ms=timeit([&]{k_scalar<1><<<blocks,threads>>>(out,1.0001f,iters);}); printf("scalar %-12s %.3f ms  %.1f G elem-updates/s\n", sn[1], ms, lanes/(ms*1e-3)/1e9);
  ms=timeit([&]{k_scalar<2><<<blocks,threads>>>(out,1.0001f,iters);}); printf("scalar %-12s %.3f ms  %.1f G elem-updates/s\n", sn[2], ms, lanes/(ms*1e-3)/1e9);
  ms=timeit([&]{k_scalar<3><<<blocks,threads>>>(out,1.0001f,iters);}); printf("scalar %-12s %.3f ms  %.1f G elem-updates/s\n", sn[3], ms, lanes/(ms*1e-3)/1e9);
  const char* pn[5]={"FFMA2+FADD2","FFMA2","FADD2","FMUL2+2FADD","FMUL2"};
  ms=timeit([&]{k_packed<0><<<blocks,threads>>>(out,1.0001f,iters,nz);}); printf("packed %-12s %.3f ms  %.1f G elem-updates/s\n", pn[0], ms, 2*lanes/(ms*1e-3)/1e9);
  ms=timeit([&]{k_packed<1><<<blocks,threads>>>(out,1.0001f,iters,nz);}); printf("packed %-12s %.3f ms  %.1f G elem-updates/s\n", pn[1], ms, 2*lanes/(ms*1e-3)/1e9);
  ms=timeit([&]{k_packed<2><<<blocks,threads>>>(out,1.0001f,iters,nz);}); printf("packed %-12s %.3f ms  %.1f G elem-updates/s\n", pn[2], ms, 2*lanes/(ms*1e-3)/1e9);
  ms=timeit([&]{k_packed<3><<<blocks,threads>>>(out,1.0001f,iters,nz);}); printf("packed %-12s %.3f ms  %.1f G elem-updates/s\n", pn[3], ms, 2*lanes/(ms*1e-3)/1e9);
  ms=timeit([&]{k_packed<4><<<blocks,threads>>>(out,1.0001f,iters,nz);}); printf("packed %-12s %.3f ms  %.1f G elem-updates/s\n", pn[4], ms, 2*lanes/(ms*1e-3)/1e9);
  printf("reference: 148 SMs x 128 lanes x 1.965 GHz = %.1f G lane-instr/s\n", sms*128*1.965);
  return 0;
}
