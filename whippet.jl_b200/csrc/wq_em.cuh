// EM stages (transcript TPM EM, PSI EM).  Filled in by the quant milestone.
#pragma once
#include <string>
#include "wq_host_quant.h"

static inline std::string wq_run_em_tpm(const WqHostIndex&, const WqIsoIndex&, WqHostQuant&, const wq_em_param&,
                                        cudaStream_t, double*, double*, double*, int64_t*) {
    return "transcript EM is not built yet";
}
static inline std::string wq_host_assign_ambig(const WqHostIndex&, const WqIsoIndex&, WqHostQuant&, wq_values*) {
    return "assign_ambig is not built yet";
}
static inline std::string wq_run_em_psi(wq_psi_batch&, cudaStream_t) { return "PSI EM is not built yet"; }
