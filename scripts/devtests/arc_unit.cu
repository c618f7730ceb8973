// dev test: device fast_arc_value vs a plain host loop on random 72-pitch patches
#include "../../orb_slam2_ros2_b200/csrc/orbx_kernels.cu"
#include <cstdlib>
#include <vector>
using namespace orbx;
__global__ void arc_k(const uint8_t* pats, int n, int* out){ int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n) out[i]=fast_arc_value(pats+(size_t)i*72*8+3*72+3); }
static int host_arc(const uint8_t* c){ const int dx[16]={0,1,2,3,3,3,2,1,0,-1,-2,-3,-3,-3,-2,-1}, dy[16]={3,3,2,1,0,-1,-2,-3,-3,-3,-2,-1,0,1,2,3}; int d[16]; for(int k=0;k<16;++k) d[k]=c[0]-c[dy[k]*72+dx[k]]; int best=-256; for(int s=0;s<16;++s){int mn=256,mx=-256; for(int i=0;i<9;++i){int q=d[(s+i)&15]; mn=q<mn?q:mn; mx=q>mx?q:mx;} best=mn>best?mn:best; best=-mx>best?-mx:best;} return best; }
int main(){ int n=100000; std::vector<uint8_t> h((size_t)n*72*8); for(auto&v:h) v=(rand()%3==0)?rand()&255:100+(rand()&31); uint8_t* d; int* o; cudaMalloc(&d,h.size()); cudaMalloc(&o,n*4); cudaMemcpy(d,h.data(),h.size(),cudaMemcpyHostToDevice); arc_k<<<(n+127)/128,128>>>(d,n,o); std::vector<int> r(n); cudaMemcpy(r.data(),o,n*4,cudaMemcpyDeviceToHost); int bad=0; for(int i=0;i<n;++i){int e=host_arc(h.data()+(size_t)i*72*8+3*72+3); if(e!=r[i]){ if(bad<5) printf("i=%d dev=%d host=%d\n",i,r[i],e); ++bad;}} printf("bad=%d of %d err=%s\n",bad,n,cudaGetErrorString(cudaGetLastError())); }
