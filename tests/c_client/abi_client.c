/* A plain C client of include/affpredict.h + include/affpredict_host.h (no Python, no C++): what a
 * maintainer's translation unit sees.  Built and run by tests/test_abi.py (CPU tier: everything up to the
 * first device call) and tests/test_gpu_parity.py (GPU tier: `--predict`).
 *   abi_client <scene.affscene> "<command>" [--predict]
 * prints one line per candidate: idx success code first_fail_step jl_cost+coll_cost */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "affpredict.h"
#include "affpredict_host.h"

int main(int argc, char** argv)
{
  if (argc < 3) { fprintf(stderr, "usage: %s scene command [--predict]\n", argv[0]); return 2; }
  const int predict = argc > 3 && strcmp(argv[3], "--predict") == 0;
  if (aff_abi_version() != AFF_ABI_VERSION) { fprintf(stderr, "header / library version mismatch\n"); return 1; }
  aff_opts opts;
  aff_default_opts(&opts);
  if (opts.dt != 0.01 || opts.n_after_steps != 500 || opts.early_exit != 0) { fprintf(stderr, "unexpected defaults\n"); return 1; }

  aff_host_scene* scene = aff_host_scene_load(argv[1]);
  if (!scene) { fprintf(stderr, "scene: %s\n", aff_host_last_error()); return 1; }
  aff_host_batch* batch = aff_host_batch_create();
  const int n = aff_host_batch_add_action(batch, scene, argv[2], 1.0, opts.dt);
  if (n <= 0) { fprintf(stderr, "action: %s\n", aff_host_last_error()); return 1; }
  printf("bodies %d ik_joints %d candidates %d tasks %zu constraints %zu\n", aff_host_num_bodies(scene),
         aff_host_num_ik_joints(scene), n, aff_host_batch_num_tasks(batch), aff_host_batch_num_constraints(batch));

  if (!predict) {
    /* the raw ABI with the shim's tables: without a device this must fail loudly, never compute on the host */
    aff_scene* dev = NULL;
    const int rc = aff_scene_create(aff_host_scene_desc(scene), 0, &dev);
    if (rc == AFF_OK) {
      printf("device present\n");
      aff_scene_destroy(dev);
    } else {
      printf("no device: status %d (%s)\n", rc, aff_last_error());
      if (rc != AFF_ERR_CUDA) return 1;
    }
  } else {
    aff_result* res = (aff_result*)calloc((size_t)n, sizeof(aff_result));
    char* msgs = (char*)calloc((size_t)n, 512);
    const int m = aff_host_predict(scene, batch, opts.dt, /*sort=*/1, /*device=*/0, res, msgs, 512);
    if (m != n) { fprintf(stderr, "predict: %s / %s\n", aff_host_last_error(), aff_last_error()); return 1; }
    for (int i = 0; i < n; ++i)
      printf("%d %d %d %d %.17g | %.60s\n", res[i].idx, res[i].success, res[i].code, res[i].first_fail_step,
             res[i].jl_cost + res[i].coll_cost, msgs + 512 * i);
    /* the ranking the shim applied is the one of aff_rank_results */
    int32_t* order = (int32_t*)calloc((size_t)n, sizeof(int32_t));
    aff_rank_results(res, (size_t)n, order);
    for (int i = 0; i < n; ++i) if (order[i] != i) { fprintf(stderr, "results are not ranked\n"); return 1; }
    free(order); free(msgs); free(res);
  }
  aff_host_batch_destroy(batch);
  aff_host_scene_destroy(scene);
  return 0;
}
