#include "vecsim_b200.h"
#include <stdio.h>
static int submit(void *q, void *ctx, AsyncJob **jobs, JobCallback *cbs, size_t n) { (void)q; (void)ctx; (void)jobs; (void)cbs; return (int)n * 0; }
int main(void) {
    VecSimParams h = {.algo = VecSimAlgo_HNSWLIB, .algoParams.hnswParams = {.type = VecSimType_FLOAT32, .dim = 4, .metric = VecSimMetric_L2, .M = 8}};
    VecSimParams t = {.algo = VecSimAlgo_TIERED,
                      .algoParams.tieredParams = {.jobQueue = 0, .jobQueueCtx = 0, .submitCb = submit, .flatBufferLimit = 16,
                                                  .primaryIndexParams = &h, .specificParams.tieredHnswParams = {.swapJobThreshold = 0}}};
    printf("%zu %zu %zu\n", sizeof(VecSimParams), sizeof(TieredIndexParams), sizeof(VecSimIndexDebugInfo));
    VecSimIndex *idx = VecSimIndex_New(&t); /* NULL without a GPU: fails loudly, no CPU fallback */
    printf("index %p (%s)\n", (void *)idx, idx ? "ok" : VecSimGPU_LastError());
    if (idx) VecSimIndex_Free(idx);
    return 0;
}
