// Box calibration: FP64 DFMA and DMMA.8x8x4 peak on the B200 under test.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
// Prints one JSON object. Not part of the product path.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

template<int ILP>
__global__ void dfma_kernel(double* out, int iters, double seed){
  double a[ILP]; double b = seed, c = 1.0 - seed*1e-9;
#pragma unroll
  for(int i=0;i<ILP;i++) a[i] = seed + i + threadIdx.x;
  for(int it=0; it<iters; ++it){
#pragma unroll
    for(int i=0;i<ILP;i++) a[i] = fma(a[i], c, b);
  }
  double s=0;
#pragma unroll
  for(int i=0;i<ILP;i++) s+=a[i];
  if(s==123.456) out[0]=s;
}

__device__ __forceinline__ void dmma(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}

template<int ILP>
__global__ void dmma_kernel(double* out, int iters, double seed){
  double c0[ILP], c1[ILP];
  double a = seed*1e-3 + threadIdx.x*1e-6, b = 1e-3;
#pragma unroll
  for(int i=0;i<ILP;i++){ c0[i]=i; c1[i]=-i; }
  for(int it=0; it<iters; ++it){
#pragma unroll
    for(int i=0;i<ILP;i++) dmma(c0[i],c1[i],a,b);
  }
  double s=0;
#pragma unroll
  for(int i=0;i<ILP;i++) s+=c0[i]+c1[i];
  if(s==123.456) out[0]=s;
}

// DMMA fed from shared memory each k-step: warp tile (TM*8) x (TN*8), operands re-read from smem every k4 step.
template<int TM,int TN>
__global__ void dmma_smem_kernel(double* out, int iters, double seed){
  extern __shared__ double sm[];
  const int ld = 36;               // 32 + 4 pad
  double* As = sm;                  // [TM*8*warps?] shared by all warps: [64][ld]
  double* Bs = sm + 64*ld;          // [32][ld+32]
  for(int i=threadIdx.x;i<64*ld+32*72;i+=blockDim.x) sm[i]=seed*1e-3+i*1e-7;
  __syncthreads();
  int lane=threadIdx.x&31, g=lane>>2, t=lane&3;
  double c0[TM][TN], c1[TM][TN];
#pragma unroll
  for(int i=0;i<TM;i++)
#pragma unroll
    for(int j=0;j<TN;j++){c0[i][j]=0;c1[i][j]=0;}
  for(int it=0; it<iters; ++it){
#pragma unroll
    for(int k=0;k<8;k++){           // 8 k4 steps = k 32
      double a[TM], b[TN];
#pragma unroll
      for(int i=0;i<TM;i++) a[i]=As[(i*8+g)*ld + k*4+t];
#pragma unroll
      for(int j=0;j<TN;j++) b[j]=Bs[(k*4+t)*72 + j*8+g];
#pragma unroll
      for(int i=0;i<TM;i++)
#pragma unroll
        for(int j=0;j<TN;j++) dmma(c0[i][j],c1[i][j],a[i],b[j]);
    }
  }
  double s=0;
#pragma unroll
  for(int i=0;i<TM;i++)
#pragma unroll
    for(int j=0;j<TN;j++) s+=c0[i][j]+c1[i][j];
  if(s==123.456) out[0]=s;
}

// Mixed issue: every warp carries NM independent DMMA chains and NF independent DFMA chains, interleaved in program order.
// If DMMA and DFMA ran on separate pipes the loop would take max(t_dmma, t_dfma); if they share the FP64 pipe, the sum.
template<int NM,int NF>
__global__ void mixed_kernel(double* out, int iters, double seed){
  double c0[NM>0?NM:1], c1[NM>0?NM:1], f[NF>0?NF:1];
  double a = seed*1e-3 + threadIdx.x*1e-6, b = 1e-3, c = 1.0 - seed*1e-9;
#pragma unroll
  for(int i=0;i<NM;i++){ c0[i]=i; c1[i]=-i; }
#pragma unroll
  for(int i=0;i<NF;i++) f[i]=seed+i+threadIdx.x;
  for(int it=0; it<iters; ++it){
#pragma unroll
    for(int i=0;i<(NM>NF?NM:NF);i++){
      if(i<NM) dmma(c0[i],c1[i],a,b);
      if(i<NF) f[i]=fma(f[i],c,b);
    }
  }
  double s=0;
#pragma unroll
  for(int i=0;i<NM;i++) s+=c0[i]+c1[i];
#pragma unroll
  for(int i=0;i<NF;i++) s+=f[i];
  if(s==123.456) out[0]=s;
}
// Same question with warp specialisation: even warps issue only DMMA, odd warps only DFMA.
template<int ILP>
__global__ void split_kernel(double* out, int iters, double seed){
  double c0[ILP], c1[ILP];
  double a = seed*1e-3 + threadIdx.x*1e-6, b = 1e-3, c = 1.0 - seed*1e-9;
#pragma unroll
  for(int i=0;i<ILP;i++){ c0[i]=i+threadIdx.x; c1[i]=-i; }
  if((threadIdx.x>>5)&1){
    for(int it=0; it<iters; ++it){
#pragma unroll
      for(int i=0;i<ILP;i++){ c0[i]=fma(c0[i],c,b); c1[i]=fma(c1[i],c,b); }   // 2*ILP DFMA per iteration
    }
  } else {
    for(int it=0; it<iters; ++it){
#pragma unroll
      for(int i=0;i<ILP;i++) dmma(c0[i],c1[i],a,b);
    }
  }
  double s=0;
#pragma unroll
  for(int i=0;i<ILP;i++) s+=c0[i]+c1[i];
  if(s==123.456) out[0]=s;
}

template<typename F> float time_ms(F f, int reps=5){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){ CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms; }
  return best;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out,8));
  printf("{\"gpu\":\"%s\",\"sms\":%d,\n",p.name,sms);
  const int iters=20000;
  // DFMA
  printf(" \"dfma\":[");
  {
    int cfgs[][2]={{1,128},{2,128},{4,128},{4,256},{8,256},{2,512},{4,512}}; // ctas/SM, threads
    bool first=true;
    for(auto&c:cfgs){
      int grid=sms*c[0], th=c[1];
      float ms=time_ms([&]{dfma_kernel<16><<<grid,th>>>(out,iters,1.0);});
      double fl=2.0*16*iters*(double)grid*th; 
      printf("%s{\"ctas_per_sm\":%d,\"threads\":%d,\"tflops\":%.2f}",first?"":",",c[0],th,fl/ms/1e9); first=false;
    }
  }
  printf("],\n \"dmma\":[");
  {
    bool first=true;
    int warps_list[]={4,8,16,32};
    for(int w:warps_list){
      int th= w>=8?256:128; int cps = w*32/th;
      int grid=sms*cps;
      float m1=time_ms([&]{dmma_kernel<1><<<grid,th>>>(out,iters,1.0);});
      float m2=time_ms([&]{dmma_kernel<2><<<grid,th>>>(out,iters,1.0);});
      float m4=time_ms([&]{dmma_kernel<4><<<grid,th>>>(out,iters,1.0);});
      float m8=time_ms([&]{dmma_kernel<8><<<grid,th>>>(out,iters,1.0);});
      float m16=time_ms([&]{dmma_kernel<16><<<grid,th>>>(out,iters,1.0);});
      double base=2.0*256*iters*(double)grid*(th/32);
      printf("%s{\"warps_per_sm\":%d,\"ilp1\":%.2f,\"ilp2\":%.2f,\"ilp4\":%.2f,\"ilp8\":%.2f,\"ilp16\":%.2f}",first?"":",",w,
        base*1/m1/1e9,base*2/m2/1e9,base*4/m4/1e9,base*8/m8/1e9,base*16/m16/1e9); first=false;
    }
  }
  printf("],\n");
  // DMMA latency: one warp, one chain
  {
    float ms=time_ms([&]{dmma_kernel<1><<<1,32>>>(out,iters,1.0);});
    printf(" \"dmma_dep_chain_ns_per_instr\":%.2f,\n", ms*1e6/iters);
    float ms2=time_ms([&]{dfma_kernel<1><<<1,32>>>(out,iters,1.0);});
    printf(" \"dfma_dep_chain_ns_per_instr\":%.2f,\n", ms2*1e6/iters);
  }
  // DMMA + DFMA mixed in one instruction stream / in neighbouring warps: do they co-issue?
  {
    int grid=sms*2, th=256;   // 16 warps per SM
    double warps=(double)grid*(th/32);
    float tm=time_ms([&]{mixed_kernel<8,0><<<grid,th>>>(out,iters,1.0);});
    float tf=time_ms([&]{mixed_kernel<0,8><<<grid,th>>>(out,iters,1.0);});
    float tx=time_ms([&]{mixed_kernel<8,8><<<grid,th>>>(out,iters,1.0);});
    float tx2=time_ms([&]{mixed_kernel<8,2><<<grid,th>>>(out,iters,1.0);});
    double fm=2.0*256*8*iters*warps, ff=2.0*32*8*iters*warps;
    printf(" \"mixed_same_warp\":{\"dmma_only_ms\":%.3f,\"dfma_only_ms\":%.3f,\"both_ms\":%.3f,\"dmma_only_tflops\":%.2f,\"dfma_only_tflops\":%.2f,"
           "\"both_total_tflops\":%.2f,\"dmma8_dfma2_ms\":%.3f,\"dmma8_dfma2_total_tflops\":%.2f,\"co_issue\":%s},\n",
           tm,tf,tx,fm/tm/1e9,ff/tf/1e9,(fm+ff)/tx/1e9,tx2,(fm+ff/4)/tx2/1e9, tx < 0.9f*(tm+tf) ? "true":"false");
    float ts=time_ms([&]{split_kernel<8><<<grid,th>>>(out,iters,1.0);});
    double fsm=2.0*256*8*iters*warps/2, fsf=2.0*32*16*iters*warps/2;
    printf(" \"mixed_warp_specialised\":{\"ms\":%.3f,\"dmma_tflops\":%.2f,\"dfma_tflops\":%.2f,\"total_tflops\":%.2f},\n",ts,fsm/ts/1e9,fsf/ts/1e9,(fsm+fsf)/ts/1e9);
  }
  // smem-fed DMMA
  printf(" \"dmma_smem\":[");
  {
    size_t smem=(64*36+32*72)*8;
    bool first=true;
    int wl[]={4,8,16};
    for(int w:wl){
      int th=w*32; if(th>512){th=512;} int cps=w*32/th; int grid=sms*cps; int it2=2000;
      float a=time_ms([&]{dmma_smem_kernel<2,2><<<grid,th,smem>>>(out,it2,1.0);});
      float b=time_ms([&]{dmma_smem_kernel<4,4><<<grid,th,smem>>>(out,it2,1.0);});
      float c=time_ms([&]{dmma_smem_kernel<4,2><<<grid,th,smem>>>(out,it2,1.0);});
      float d=time_ms([&]{dmma_smem_kernel<4,8><<<grid,th,smem>>>(out,it2,1.0);});
      double base=2.0*256*8*it2*(double)grid*(th/32);
      printf("%s{\"warps_per_sm\":%d,\"t2x2\":%.2f,\"t4x4\":%.2f,\"t4x2\":%.2f,\"t4x8\":%.2f}",first?"":",",w,base*4/a/1e9,base*16/b/1e9,base*8/c/1e9,base*32/d/1e9); first=false;
    }
  }
  printf("]}\n");
  return 0;
}
