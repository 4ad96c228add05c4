#include "engine.h"
namespace cri {
int plan_hca_decode(cri_ctx*, cri_job*) { return ERR_UNSUPPORTED; }
int plan_hca_crypt(cri_ctx*, cri_job*) { return ERR_UNSUPPORTED; }
int plan_hca_encode(cri_ctx*, cri_job*) { return ERR_UNSUPPORTED; }
int upload_hca_tables(cri_ctx*, cri_job*) { return OK; }
int run_hca(cri_ctx*, cri_job*, bool*) { return ERR_UNSUPPORTED; }
void free_hca_tables(cri_job*) {}
}  // namespace cri
