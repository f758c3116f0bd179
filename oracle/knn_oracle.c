/* CPU restatement of the nearest-neighbour label transfer on stage-1's chunk merge — TEST INFRASTRUCTURE ONLY.
 *
 * Reference: infinicube/voxelgen/utils/color_util.py:52-60 (semantic_from_points) ->
 * infinicube/voxelgen/ext/common/knn.cu:15-50 (knn_query_fast: FLANN-style KD-tree on the GPU,
 * kdtree_cuda.cu:1070-1094 knnSearch), called from infinicube/voxelgen/utils/extrap_util.py:233-276 and
 * infinicube/inference/voxel_generation_single_chunk.py:280.  SURVEY.md §8(f) N4.
 *
 * The exact nearest neighbour under squared L2 is unique except for ties, so the oracle is the brute-force
 * definition: d2 = ((qx-px)^2 + (qy-py)^2) + (qz-pz)^2 in fp32 without FMA contraction (compile with
 * -ffp-contract=off), arg-min with ties -> smallest reference index.  PARITY: the reference's KD-tree is CUDA
 * code inside a torch extension and cannot run in this container (no GPU) nor be copied; the oracle is pinned
 * instead against scipy.spatial.cKDTree and against the reference's own small-cloud branch (torch.cdist + topk,
 * knn.cu:23-28, restated with the same torch calls) in tests/test_oracle_knn.py.  Which of several equidistant points the
 * reference's tree returns is unspecified by its code (heap order) -> tie-break parity unpinned.
 */
#include <math.h>
#include <stddef.h>

void ko_nn1(const float* ref, long long m, int ref_stride, const float* q, long long n, int q_stride, int* idx, float* d2) {
  for (long long i = 0; i < n; ++i) {
    const float qx = q[i * q_stride], qy = q[i * q_stride + 1], qz = q[i * q_stride + 2];
    float best = INFINITY;
    int bi = -1;
    for (long long j = 0; j < m; ++j) {
      const float dx = qx - ref[j * ref_stride], dy = qy - ref[j * ref_stride + 1], dz = qz - ref[j * ref_stride + 2];
      const float d = (dx * dx + dy * dy) + dz * dz;
      if (d < best) {
        best = d;
        bi = (int)j;
      }
    }
    idx[i] = bi;
    d2[i] = best;
  }
}
