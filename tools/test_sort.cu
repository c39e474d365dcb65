// Developer test: fast_sort (register/shuffle bitonic) of nms.cu in isolation.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../multi-task-learning-object-detection-semantic-segmentation_b200/csrc/nms.cu"
bool ssdx_debug_sync() { return false; }
void ssdx_set_error(const char*, ...) {}
int ssdx_sm_count(int) { return 148; }
__global__ void k(unsigned long long* data, int n) {
  __shared__ unsigned long long s_list[1024];
  __shared__ unsigned long long s_tmp[1024];
  for (int i = threadIdx.x; i < n; i += 256) s_list[i] = data[i];
  __syncthreads();
  fast_sort(s_list, s_tmp, n, threadIdx.x);
  for (int i = threadIdx.x; i < 1024; i += 256) data[i] = s_list[i];
}
int main() {
  for (int n : {0, 1, 5, 64, 300, 512, 513, 900, 1024}) {
    std::vector<unsigned long long> h(1024);
    for (auto& x : h) x = ((unsigned long long)rand() << 32) ^ rand() | 1ull;
    unsigned long long* d;
    cudaMalloc(&d, 1024 * 8);
    cudaMemcpy(d, h.data(), 1024 * 8, cudaMemcpyHostToDevice);
    k<<<1, 256>>>(d, n);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned long long> o(1024);
    cudaMemcpy(o.data(), d, 1024 * 8, cudaMemcpyDeviceToHost);
    std::vector<unsigned long long> ref(h.begin(), h.begin() + n);
    std::sort(ref.rbegin(), ref.rend());
    bool ok = e == cudaSuccess;
    for (int i = 0; i < n && ok; ++i) ok = o[i] == ref[i];
    printf("n=%d %s %s\n", n, cudaGetErrorString(e), ok ? "OK" : "MISMATCH");
    cudaFree(d);
  }
  return 0;
}
