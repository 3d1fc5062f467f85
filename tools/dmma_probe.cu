// Probe: what costs the DMMA GEMM mainloop its last 25%?  64x64 CTA tile, 4 warps (TM=TN=4), k-tiles of BK.
//   mode 0: LDS + DMMA only (operands static in smem)        mode 1: + __syncthreads per k-tile
//   mode 2: + 3-stage cp.async ring from global (full NT mainloop)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)
__device__ __forceinline__ void dmma(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
__device__ __forceinline__ void cp16(void* s, const void* g){ unsigned a=(unsigned)__cvta_generic_to_shared(s); asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"::"r"(a),"l"(g)); }

template<int BK,int MODE,int STAGES>
__global__ void __launch_bounds__(128) probe(const double* __restrict__ A, const double* __restrict__ B, double* out, int ktiles, long ld){
  constexpr int LDA=BK+4, ST=64*LDA;
  extern __shared__ __align__(16) double sm[];
  double* As=sm; double* Bs=sm+STAGES*ST;
  const int tid=threadIdx.x, lane=tid&31, warp=tid>>5, g8=lane>>2, t4=lane&3, wm=warp>>1, wn=warp&1;
  for(int i=tid;i<2*STAGES*ST;i+=128) sm[i]=1e-3+i*1e-7;
  __syncthreads();
  const double* Ab=A+(long)blockIdx.x%64*64*ld; const double* Bb=B+(long)(blockIdx.x/64%64)*64*ld;
  auto load=[&](int stage,int k0){
    for(int c=tid;c<64*BK/2;c+=128){ int r=c/(BK/2), cc=c%(BK/2);
      cp16(As+stage*ST+r*LDA+cc*2, Ab+(long)r*ld+k0+cc*2);
      cp16(Bs+stage*ST+r*LDA+cc*2, Bb+(long)r*ld+k0+cc*2); } };
  double c0[4][4],c1[4][4];
#pragma unroll
  for(int i=0;i<4;i++)
#pragma unroll
    for(int j=0;j<4;j++){c0[i][j]=0;c1[i][j]=0;}
  if(MODE==2){ for(int s=0;s<STAGES-1;s++){ load(s,s*BK); asm volatile("cp.async.commit_group;"); } }
  for(int kt=0;kt<ktiles;++kt){
    if(MODE==2){ asm volatile("cp.async.wait_group %0;"::"n"(STAGES-2)); }
    if(MODE>=1) __syncthreads();
    if(MODE==2){ int nk=kt+STAGES-1; if(nk<ktiles) load(nk%STAGES,nk*BK); asm volatile("cp.async.commit_group;"); }
    const int st=kt%STAGES;
    const double* as=As+st*ST+(wm*32+g8)*LDA+t4; const double* bs=Bs+st*ST+(wn*32+g8)*LDA+t4;
#pragma unroll
    for(int kk=0;kk<BK/4;kk++){
      double a[4],b[4];
#pragma unroll
      for(int i=0;i<4;i++) a[i]=as[i*8*LDA+kk*4];
#pragma unroll
      for(int j=0;j<4;j++) b[j]=bs[j*8*LDA+kk*4];
#pragma unroll
      for(int i=0;i<4;i++)
#pragma unroll
        for(int j=0;j<4;j++) dmma(c0[i][j],c1[i][j],a[i],b[j]);
    }
  }
  double s=0;
#pragma unroll
  for(int i=0;i<4;i++)
#pragma unroll
    for(int j=0;j<4;j++) s+=c0[i][j]+c1[i][j];
  if(s==123.456) out[0]=s;
}
template<int BK,int MODE,int STAGES> void run(const double* A,const double* B,double* out,int sms,long ld,const char* name){
  size_t smem=2*STAGES*64*(BK+4)*8; auto k=probe<BK,MODE,STAGES>;
  CK(cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  int per=0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per,k,128,smem));
  for(int cps=1;cps<=per && cps<=4;cps++){
    int grid=sms*cps; int ktiles=(int)(ld/BK);
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<grid,128,smem>>>(A,B,out,ktiles,ld); CK(cudaDeviceSynchronize());
    float best=1e30f; for(int r=0;r<3;r++){ CK(cudaEventRecord(e0)); k<<<grid,128,smem>>>(A,B,out,ktiles,ld); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms; }
    double fl=2.0*64*64*(double)ktiles*BK*grid;
    printf("%-34s ctas/SM %d (max %d): %7.2f TFLOP/s\n",name,cps,per,fl/best/1e9);
  }
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  long ld=65536; double *A,*B,*out; CK(cudaMalloc(&A,64*64*ld*8)); CK(cudaMalloc(&B,64*64*ld*8)); CK(cudaMalloc(&out,8));
  CK(cudaMemset(A,0,64*64*ld*8)); CK(cudaMemset(B,0,64*64*ld*8));
  run<16,0,3>(A,B,out,sms,ld,"BK16 LDS+DMMA");
  run<16,1,3>(A,B,out,sms,ld,"BK16 +syncthreads");
  run<16,2,3>(A,B,out,sms,ld,"BK16 +cp.async ring (3 stages)");
  run<16,2,4>(A,B,out,sms,ld,"BK16 +cp.async ring (4 stages)");
  run<32,0,3>(A,B,out,sms,ld,"BK32 LDS+DMMA");
  run<32,1,3>(A,B,out,sms,ld,"BK32 +syncthreads");
  run<32,2,3>(A,B,out,sms,ld,"BK32 +cp.async ring (3 stages)");
  return 0;
}
