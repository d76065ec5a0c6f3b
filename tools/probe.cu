// Hardware/sandbox capability probe: CUDA IPC between processes, stream memOps on
// local + peer memory, cross-process release/acquire flags, host-register of stack
// memory, IPC-open while a spin kernel is resident. Run: ./probe  (forks 2 ranks)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <thread>
#include <atomic>
#include <chrono>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("[r%d] %s:%d %s -> %s\n",rank,__FILE__,__LINE__,#x,cudaGetErrorString(e)); fflush(stdout); }}while(0)
#define CU(x) do{CUresult e=(x); if(e!=CUDA_SUCCESS){const char*s; cuGetErrorString(e,&s); printf("[r%d] %s:%d %s -> %s\n",rank,__FILE__,__LINE__,#x,s); fflush(stdout);}}while(0)
static int rank;
struct Shared { cudaIpcMemHandle_t h[2]; volatile int ready[2]; volatile int phase[2]; double t[2]; };
static void sbar(Shared*s,int p){ __sync_synchronize(); s->phase[rank]=p; while(s->phase[1-rank]<p) usleep(10); __sync_synchronize(); }
__global__ void spin_acq(unsigned long long* f, unsigned long long v, long long* cyc){
  long long t0=clock64(); unsigned long long x;
  do { asm volatile("ld.acquire.sys.global.u64 %0,[%1];":"=l"(x):"l"(f)); } while(x<v);
  if(cyc) *cyc=clock64()-t0;
}
__global__ void st_rel(unsigned long long* f, unsigned long long v){
  asm volatile("st.release.sys.global.u64 [%0],%1;"::"l"(f),"l"(v));
}
__global__ void copyk(int4* d,const int4* s,size_t n){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st) d[i]=s[i]; }
int main(){
  Shared* s=(Shared*)mmap(0,sizeof(Shared),PROT_READ|PROT_WRITE,MAP_SHARED|MAP_ANONYMOUS,-1,0);
  memset(s,0,sizeof(*s));
  pid_t pid=fork(); rank = pid==0?1:0;
  int ndev=0; CK(cudaGetDeviceCount(&ndev)); int dev=rank%ndev; CK(cudaSetDevice(dev));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,dev));
  if(rank==0){ printf("ndev=%d name=%s sms=%d\n",ndev,pr.name,pr.multiProcessorCount);
    int v; cudaDeviceGetAttribute(&v,cudaDevAttrPageableMemoryAccess,dev); printf("pageableMemoryAccess=%d\n",v);
    cudaDeviceGetAttribute(&v,cudaDevAttrHostRegisterSupported,dev); printf("hostRegisterSupported=%d\n",v);
    CUdevice cd; cuDeviceGet(&cd,dev); cuDeviceGetAttribute(&v,CU_DEVICE_ATTRIBUTE_CAN_USE_64_BIT_STREAM_MEM_OPS,cd); printf("memops64=%d\n",v);
    cuDeviceGetAttribute(&v,CU_DEVICE_ATTRIBUTE_CAN_FLUSH_REMOTE_WRITES,cd); printf("flushRemoteWrites=%d\n",v);
    cuDeviceGetAttribute(&v,CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED,cd); printf("vmmPosixFd=%d\n",v);
    cuDeviceGetAttribute(&v,CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED,cd); printf("multicast=%d\n",v);
    cuDeviceGetAttribute(&v,CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED,cd); printf("fabric=%d\n",v);
    if(ndev>1){int pa; cudaDeviceCanAccessPeer(&pa,0,1); printf("canAccessPeer(0,1)=%d\n",pa);}
    fflush(stdout);}
  size_t N=256u<<20;
  char* buf; CK(cudaMalloc(&buf,N+4096)); CK(cudaMemset(buf,0,N+4096));
  CK(cudaIpcGetMemHandle(&s->h[rank],buf)); sbar(s,1);
  char* peer=0; CK(cudaIpcOpenMemHandle((void**)&peer,s->h[1-rank],cudaIpcMemLazyEnablePeerAccess));
  printf("[r%d] dev=%d ipc open peer=%p\n",rank,dev,peer); fflush(stdout);
  unsigned long long* lflag=(unsigned long long*)(buf+N); unsigned long long* pflag=(unsigned long long*)(peer+N);
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st,cudaStreamNonBlocking));
  sbar(s,2);
  // Test A: memop write to PEER memory, memop wait on LOCAL memory. ping-pong 1000 iters.
  { int iters=1000; auto t0=std::chrono::steady_clock::now();
    for(int i=1;i<=iters;i++){
      if(rank==0){ CU(cuStreamWriteValue64(st,(CUdeviceptr)pflag,(cuuint64_t)i,0)); CU(cuStreamWaitValue64(st,(CUdeviceptr)lflag,(cuuint64_t)i,CU_STREAM_WAIT_VALUE_GEQ)); }
      else       { CU(cuStreamWaitValue64(st,(CUdeviceptr)lflag,(cuuint64_t)i,CU_STREAM_WAIT_VALUE_GEQ)); CU(cuStreamWriteValue64(st,(CUdeviceptr)pflag,(cuuint64_t)i,0)); }
    }
    CK(cudaStreamSynchronize(st)); auto t1=std::chrono::steady_clock::now();
    printf("[r%d] memop pingpong (write peer / wait local): %.2f us round trip\n",rank,std::chrono::duration<double,std::micro>(t1-t0).count()/iters); fflush(stdout);
  }
  sbar(s,3);
  // Test B: kernel st.release.sys to peer; kernel spin ld.acquire local. ping-pong
  { CK(cudaMemset(lflag,0,8)); CK(cudaDeviceSynchronize()); sbar(s,4);
    int iters=200; auto t0=std::chrono::steady_clock::now();
    for(int i=1;i<=iters;i++){
      if(rank==0){ st_rel<<<1,1,0,st>>>(pflag,i); spin_acq<<<1,1,0,st>>>(lflag,i,0); }
      else       { spin_acq<<<1,1,0,st>>>(lflag,i,0); st_rel<<<1,1,0,st>>>(pflag,i); }
    }
    CK(cudaStreamSynchronize(st)); auto t1=std::chrono::steady_clock::now();
    printf("[r%d] kernel pingpong: %.2f us round trip\n",rank,std::chrono::duration<double,std::micro>(t1-t0).count()/iters); fflush(stdout);
  }
  sbar(s,5);
  // Test C: push bandwidth to peer with a simple copy kernel + cudaMemcpyPeer-like D2D
  { cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    char* src; CK(cudaMalloc(&src,N)); CK(cudaMemset(src,1,N));
    for(int rep=0;rep<2;rep++){
      if(rank==0){ cudaEventRecord(e0,st); for(int i=0;i<5;i++) copyk<<<592,512,0,st>>>((int4*)peer,(int4*)src,N/16); cudaEventRecord(e1,st); CK(cudaStreamSynchronize(st)); float ms; cudaEventElapsedTime(&ms,e0,e1); printf("[r0] kernel push %zu MiB: %.1f GB/s\n",N>>20, 5.0*N/ms/1e6);
        cudaEventRecord(e0,st); for(int i=0;i<5;i++) cudaMemcpyAsync(peer,src,N,cudaMemcpyDeviceToDevice,st); cudaEventRecord(e1,st); CK(cudaStreamSynchronize(st)); cudaEventElapsedTime(&ms,e0,e1); printf("[r0] memcpyAsync push: %.1f GB/s\n", 5.0*N/ms/1e6);
        cudaEventRecord(e0,st); for(int i=0;i<5;i++) copyk<<<592,512,0,st>>>((int4*)src,(int4*)peer,N/16); cudaEventRecord(e1,st); CK(cudaStreamSynchronize(st)); cudaEventElapsedTime(&ms,e0,e1); printf("[r0] kernel pull: %.1f GB/s\n", 5.0*N/ms/1e6); fflush(stdout);}
      sbar(s,6+rep);
    }
  }
  sbar(s,10);
  // Test D: host register stack memory + mapped access
  { int stackv=rank+41; cudaError_t e=cudaHostRegister(&stackv,sizeof(int),cudaHostRegisterMapped); printf("[r%d] hostRegister(stack) -> %s\n",rank,cudaGetErrorString(e));
    if(e==cudaSuccess){ int* dp; CK(cudaHostGetDevicePointer((void**)&dp,&stackv,0)); CK(cudaMemcpyAsync(buf,dp,4,cudaMemcpyDeviceToDevice,st)); CK(cudaStreamSynchronize(st)); int back=0; CK(cudaMemcpy(&back,buf,4,cudaMemcpyDeviceToHost)); printf("[r%d] read stack via mapped ptr = %d\n",rank,back); cudaHostUnregister(&stackv);} fflush(stdout);
  }
  sbar(s,11);
  // Test E: while a spin kernel is resident on stream st, do cudaMalloc + IpcGet/Open + HostRegister from another thread
  { CK(cudaMemset(lflag,0,8)); CK(cudaDeviceSynchronize());
    spin_acq<<<1,1,0,st>>>(lflag,1,0);
    std::atomic<int> done{0};
    std::thread th([&]{ cudaSetDevice(dev); auto t0=std::chrono::steady_clock::now(); char* b2; cudaError_t e1=cudaMalloc(&b2,1<<20); cudaIpcMemHandle_t h; cudaError_t e2=cudaIpcGetMemHandle(&h,b2);
       static int hv[1024]; cudaError_t e3=cudaHostRegister(hv,4096,cudaHostRegisterMapped);
       char* p2=0; cudaError_t e4=cudaIpcOpenMemHandle((void**)&p2,s->h[1-rank],cudaIpcMemLazyEnablePeerAccess); // already open: expect error or same ptr
       auto t1=std::chrono::steady_clock::now();
       printf("[r%d] under-spin: malloc=%s ipcget=%s hostreg=%s ipcopen(again)=%s took %.1f us\n",rank,cudaGetErrorName(e1),cudaGetErrorName(e2),cudaGetErrorName(e3),cudaGetErrorName(e4),std::chrono::duration<double,std::micro>(t1-t0).count()); fflush(stdout); done=1;});
    for(int i=0;i<300 && !done;i++) usleep(10000);
    printf("[r%d] under-spin thread done=%d (0 => CUDA calls block behind a spinning kernel)\n",rank,(int)done); fflush(stdout);
    unsigned long long one=1; cudaStream_t s2; cudaStreamCreateWithFlags(&s2,cudaStreamNonBlocking); CK(cudaMemcpyAsync(lflag,&one,8,cudaMemcpyHostToDevice,s2)); CK(cudaStreamSynchronize(st)); th.join();
  }
  sbar(s,12);
  if(rank==0){ int stt; wait(&stt); printf("probe done\n"); }
  return 0;
}
