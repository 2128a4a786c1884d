/* Plain-C client of include/gpb200.h: proves the boundary is a C ABI (the header compiles as C99, every declared
 * entry point resolves with dlsym, argument validation returns the documented status codes before any CUDA call).
 * Built and run by tests/test_abi.py; needs no GPU. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "gpb200.h"

#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) { fprintf(stderr, "FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } \
  } while (0)

typedef int (*version_fn)(void);
typedef int (*block_fn)(void);
typedef const char* (*err_fn)(void);
typedef size_t (*potri_ws_fn)(int);
typedef int (*kern_fwd_fn)(int, const double*, int, long, const double*, int, long, int, const double*, int,
                           const double*, const double*, int, double*, long, void*);
typedef int (*gemm_fn)(int, int, int, int, double, const double*, long, const double*, long, double, double*, long,
                       int, void*);
typedef int (*potrf_fn)(double*, int, long, double*, int*, void*);

int main(int argc, char** argv) {
  static const char* names[] = {
      "gpb_version", "gpb_last_error", "gpb_block_size", "gpb_launch_count", "gpb_reset_launch_count",
      "gpb_dmma_issue_probe", "gpb_kern_fwd",
      "gpb_kern_bwd_workspace_bytes", "gpb_kern_bwd", "gpb_kern_bwd_mul", "gpb_kern_sop_fwd", "gpb_linear_kdiag",
      "gpb_potrf_lower", "gpb_tri_diag_inverse", "gpb_potri_workspace_bytes", "gpb_potri_lower", "gpb_trtri_upper",
      "gpb_potri_assemble", "gpb_tri_zero_upper", "gpb_add_diag", "gpb_trsv_workspace_bytes", "gpb_trsv_lower",
      "gpb_trsm_right_lt", "gpb_logdet_sumsq", "gpb_rowdot", "gpb_gemv_n",
      "gpb_rows_scale_add_outer", "gpb_gemv_t_workspace_bytes", "gpb_gemv_t", "gpb_gemm", "gpb_gemm_splitk",
      "gpb_gemm_ozaki_nt", "gpb_ozaki_config",
      "gpb_gpr_grad_workspace_bytes", "gpb_gpr_grad", "gpb_kuf_stats_workspace_bytes", "gpb_kuf_stats_fwd",
      "gpb_kuf_stats_bwd"};
  void* lib;
  size_t i;
  version_fn version;
  block_fn block;
  err_fn last_error;
  potri_ws_fn potri_ws;
  kern_fwd_fn kern_fwd;
  gemm_fn gemm;
  potrf_fn potrf;
  double x[4] = {0.0, 1.0, 2.0, 3.0};
  CHECK(argc == 2);
  lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
  for (i = 0; i < sizeof(names) / sizeof(names[0]); ++i) {
    if (!dlsym(lib, names[i])) { fprintf(stderr, "missing symbol %s\n", names[i]); return 1; }
  }
  /* POSIX idiom for object-pointer -> function-pointer (ISO C has no such conversion) */
  *(void**)(&version) = dlsym(lib, "gpb_version");
  *(void**)(&block) = dlsym(lib, "gpb_block_size");
  *(void**)(&last_error) = dlsym(lib, "gpb_last_error");
  *(void**)(&potri_ws) = dlsym(lib, "gpb_potri_workspace_bytes");
  *(void**)(&kern_fwd) = dlsym(lib, "gpb_kern_fwd");
  *(void**)(&gemm) = dlsym(lib, "gpb_gemm");
  *(void**)(&potrf) = dlsym(lib, "gpb_potrf_lower");
  CHECK(version() >= 100);
  CHECK(block() == 128);
  CHECK(last_error() != NULL);
  CHECK(potri_ws(4096) > 0);
  /* argument validation happens before any CUDA call: these return without touching a device */
  CHECK(kern_fwd(GPB_KERN_RBF, NULL, 4, 1, NULL, 0, 0, 1, x, 1, x, NULL, GPB_FILL_FULL, NULL, 4, NULL) == GPB_ERR_BADARG);
  CHECK(kern_fwd(99, x, 4, 1, NULL, 0, 0, 1, x, 1, x, NULL, GPB_FILL_FULL, x, 4, NULL) == GPB_ERR_BADARG);
  CHECK(gemm(7, 4, 4, 4, 1.0, x, 4, x, 4, 0.0, x, 4, 0, NULL) == GPB_ERR_BADARG);
  CHECK(gemm(0, 0, 4, 4, 1.0, x, 4, x, 4, 0.0, x, 4, 0, NULL) == GPB_OK); /* empty product */
  CHECK(potrf(NULL, 4, 4, NULL, NULL, NULL) == GPB_ERR_BADARG);
  CHECK(potrf(NULL, 0, 0, NULL, NULL, NULL) == GPB_OK); /* n = 0 */
  printf("abi_check: %d symbols ok\n", (int)(sizeof(names) / sizeof(names[0])));
  dlclose(lib);
  return 0;
}
