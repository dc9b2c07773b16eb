// tcgen05 candidate pass for the kNN (placeholder until the TMEM kernel lands:
// method 2 currently routes to the exact fp64 sweep so that the ABI is stable).
#include "rf_common.cuh"

int rf_knn_exact_launch(const float* bank, long n_rows, long row_offset, const float* q, long Q, int k, int* out_idx,
                        double* out_d, void* workspace, size_t workspace_bytes, cudaStream_t s);

size_t rf_knn_tc_workspace_bytes(long Q, long n_rows, int k) { (void)Q; (void)n_rows; (void)k; return 256; }

int rf_knn_tc_launch(const float* bank, long n_rows, long row_offset, const float* q, long Q, int k, int* out_idx,
                     double* out_d, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    return rf_knn_exact_launch(bank, n_rows, row_offset, q, Q, k, out_idx, out_d, workspace, workspace_bytes, s);
}
