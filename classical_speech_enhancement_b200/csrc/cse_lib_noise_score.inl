// temporary stubs (replaced as the kernels land)
extern "C" {
size_t cse_noise_workspace_bytes(int, int, int) { return 0; }
int cse_noise_percentile(const void*, int, int, int, double, double, void*, void*, size_t, void*) { return fail(CSE_EUNSUPPORTED, "todo"); }
int cse_noise_mintrack(const void*, int, int, int, double, void*, void*, size_t, void*) { return fail(CSE_EUNSUPPORTED, "todo"); }
size_t cse_clean_cache_bytes(int, int) { return 0; }
size_t cse_clean_workspace_bytes(int, int, int) { return 0; }
int cse_prepare_clean(const void*, const void*, int, int, int, void*, void*, size_t, void*) { return fail(CSE_EUNSUPPORTED, "todo"); }
size_t cse_score_workspace_bytes(int, int, int) { return 0; }
int cse_score(const void*, const void*, int, int, int, int, const void*, const void*, int, cse_score_t*, void*, size_t, void*) { return fail(CSE_EUNSUPPORTED, "todo"); }
size_t cse_sweep_workspace_bytes(int, int, int) { return 0; }
int cse_sweep(const void*, int, const void*, const void*, int, int, int, int, int, const cse_params*, int, int, const void*, const void*, cse_score_t*, int, void*, size_t, void*) { return fail(CSE_EUNSUPPORTED, "todo"); }
}
