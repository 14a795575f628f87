// microbenchmark: issue/throughput of scalar FMUL+FADD vs packed FFMA2(-0)+FADD2 (bit-identical strict pairs) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
constexpr int NA = 8;
template<int MODE> __global__ void k(float* out, int iters, float m, float s){
  float acc[2*NA]; u64 acc2[NA];
  for(int i=0;i<2*NA;i++) acc[i]=threadIdx.x*0.001f+i;
  for(int i=0;i<NA;i++) acc2[i]=pk(acc[2*i],acc[2*i+1]);
  const u64 nz = pk(-0.0f,-0.0f), mm = pk(m,m*1.0001f), ss = pk(s,s);
  for(int it=0;it<iters;it++){
    if (MODE==0){ // scalar strict: mul + add
      #pragma unroll
      for(int i=0;i<2*NA;i++){ float p = __fmul_rn(acc[i], m); acc[i] = __fadd_rn(p, s); }
    } else if (MODE==1){ // scalar FFMA
      #pragma unroll
      for(int i=0;i<2*NA;i++) acc[i] = __fmaf_rn(acc[i], m, s);
    } else if (MODE==2){ // packed strict: FFMA2(-0) + FADD2
      #pragma unroll
      for(int i=0;i<NA;i++){ u64 p = fma2(acc2[i], mm, nz); acc2[i] = add2(p, ss); }
    } else { // packed FFMA2
      #pragma unroll
      for(int i=0;i<NA;i++) acc2[i] = fma2(acc2[i], mm, ss);
    }
  }
  float r=0; for(int i=0;i<2*NA;i++) r+=acc[i];
  for(int i=0;i<NA;i++){ float a,b; upk(acc2[i],a,b); r+=a+b; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int MODE> void run(const char* name, int warps_per_sm){
  float* out; cudaMalloc(&out, 148*1024*4*sizeof(float));
  int iters=20000; cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, warps_per_sm*32>>>(out, 100, 1.0001f, 0.5f); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<148, warps_per_sm*32>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double elems = 148.0*warps_per_sm*32*iters*2*NA; // element-updates (each = 1 mul + 1 add)
  printf("%-28s warps/SM=%2d  %.3f ms  %.1f G elem-updates/s  (%.2f per clk per SM @1.965GHz)\n", name, warps_per_sm, ms, elems/ms/1e6, elems/(ms*1e-3)/148/1.965e9);
  cudaFree(out);
}
int main(){
  for (int w : {4, 8, 16, 32}) {
    run<0>("scalar FMUL+FADD (strict)", w); run<1>("scalar FFMA", w); run<2>("packed FFMA2(-0)+FADD2", w); run<3>("packed FFMA2", w);
  }
  return 0;
}
