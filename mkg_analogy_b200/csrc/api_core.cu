// Error plumbing + the stand-alone GEMM entry point of the C ABI (include/b200mart.h).
#include "gemm.cuh"
#include "../../include/b200mart.h"
#include <stdarg.h>
#include <stdlib.h>

thread_local char g_b200mart_err[512] = "";
unsigned long long g_b200mart_launches = 0;

int b200mart_set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_b200mart_err, sizeof(g_b200mart_err), fmt, ap);
    va_end(ap);
    return code;
}

int b200mart_pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("B200MART_PDL"); on = e ? (atoi(e) != 0) : 1; }
    return on;
}

extern "C" const char* b200mart_last_error(void) { return g_b200mart_err; }

extern "C" int32_t b200mart_version(void) { return B200MART_VERSION; }

extern "C" int64_t b200mart_launch_count(void) { return (int64_t)g_b200mart_launches; }

extern "C" int32_t b200mart_gemm_bf16(const b200mart_gemm_desc* d, void* stream) {
    if (!d) return b200mart_set_err(1, "gemm: null descriptor");
    GemmOp op;
    op.M = d->M; op.N = d->N; op.K1 = d->K; op.K2 = d->K2;
    op.a_mn = d->a_mn; op.b_mn = d->b_mn;
    op.A = (const bf16*)d->A; op.lda = d->lda; op.B = (const bf16*)d->B; op.ldb = d->ldb;
    op.A2 = (const bf16*)d->A2; op.lda2 = d->lda2; op.B2 = (const bf16*)d->B2; op.ldb2 = d->ldb2;
    op.alpha = d->alpha; op.bias = d->bias; op.bias2 = d->bias2;
    op.out_preact = (bf16*)d->out_preact; op.ldp = d->ldp;
    op.act = d->act; op.save_grad = d->save_grad; op.dact_src = (const bf16*)d->dact_src; op.ldx = d->ldx; op.dact = d->dact;
    op.residual = d->residual; op.ldr = d->ldr;
    op.out_f32 = d->out_f32; op.ldo32 = d->ldo32; op.out_bf16 = (bf16*)d->out_bf16; op.ldo16 = d->ldo16;
    op.accumulate = d->accumulate; op.split_k = d->split_k > 0 ? d->split_k : 1; op.bn = d->bn;
    GemmPlan* plan = nullptr;
    B2_TRY(gemm_plan_create(op, &plan));
    int rc = gemm_plan_launch(plan, (cudaStream_t)stream);
    gemm_plan_destroy(plan);
    return rc;
}
