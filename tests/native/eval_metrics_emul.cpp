// CPU-thread emulation build of arvae_b200/csrc/eval_metrics.cu (TEST INFRASTRUCTURE ONLY): the kernels and the
// host orchestration compile unchanged against cuda_emul.h; only the argsort is replaced by std::sort.
#define ARVAE_HOST_EMULATION 1
#include "cuda_emul.h"
#include "../../arvae_b200/csrc/eval_metrics.cu"

#include <vector>

extern "C" int emul_eval_metrics(const float *codes, int64_t crs, int64_t ccs, const float *attrs, int64_t ars,
                                 int64_t acs, int64_t B, int Z, int A, double *rho, double *pval, double *corr,
                                 double *sap, double *scores) {
    std::vector<char> ws(arvae::eval_metrics_workspace_bytes(B, Z, A) + 256);
    char *base = ws.data() + (256 - (reinterpret_cast<uintptr_t>(ws.data()) & 255)) % 256;
    return arvae::run_eval_metrics(codes, crs, ccs, attrs, ars, acs, B, Z, A, rho, pval, corr, sap, scores, base,
                                   nullptr);
}
