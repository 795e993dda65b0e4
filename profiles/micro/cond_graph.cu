#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* c, cudaGraphConditionalHandle hnd) {
    int v = atomicSub(c, 1) - 1;
    if (threadIdx.x == 0) cudaGraphSetConditional(hnd, v > 0);
}
int main() {
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle hnd;
    cudaGraphConditionalHandleCreate(&hnd, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
    p.conditional.handle = hnd; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node;
    cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
    printf("add %d\n", (int)e);
    cudaGraph_t bg = p.conditional.phGraph_out[0];
    int* d; cudaMalloc(&d, 4); int five = 5; cudaMemcpy(d, &five, 4, cudaMemcpyHostToDevice);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaStreamBeginCaptureToGraph(s, bg, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
    body<<<1,1,0,s>>>(d, hnd);
    cudaStreamEndCapture(s, nullptr);
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %d\n", (int)e);
    e = cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
    int out; cudaMemcpy(&out, d, 4, cudaMemcpyDeviceToHost); printf("launch %d out %d\n", (int)e, out);
}
