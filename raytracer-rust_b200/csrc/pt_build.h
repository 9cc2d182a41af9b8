// pt_build.h — host-side flattening of a Mesh into the device layout (pt_types.h).
//
// Three steps per mesh (done once, at ptc_scene_commit):
//  1. restate the reference's BVH build (BVHNode::new, src/acceleration/bvh.rs:15-76) far enough to know, for
//     every triangle, (a) whether it sits under a zero-extent node — such nodes can never be entered
//     (Aabb::intersect, src/acceleration/aabb.rs:40) so the triangle is invisible in the reference — and (b) its
//     position in the depth-first leaf order, which decides equal-t ties (bvh.rs:142-156);
//  2. build a binned-SAH binary BVH over the LIVE triangles only;
//  3. collapse it to 8-wide nodes, assign children to octant-ordered slots and quantise child boxes to 8 bits.
#pragma once
#include <cstdint>
#include <vector>

#include "pt_types.h"

namespace pt {

struct MeshBuild {
  // input: n x 12 floats = v0, v1, v2, normal (object space), as handed to ptc_scene_add_mesh
  std::vector<float> tris;
  int64_t n = 0;

  // step 1
  std::vector<uint8_t> dead;   // n
  std::vector<int32_t> order;  // n: position in the reference's DFS leaf order
  int64_t ref_nodes = 0, ref_leaves = 0, live = 0;
  int32_t ref_depth = 0;
  bool ref_done = false;  // step 1 already done elsewhere (pt_build_dev.cu): build_mesh starts at step 2

  // steps 2-3
  std::vector<Node8> nodes;
  std::vector<Tri48> tri48;
  std::vector<float4> normals;  // n: Triangle.normal by DFS position (`order`), original index bit-cast in .w
  int32_t wide_depth = 0;
  float root_lo[3] = {1, 1, 1}, root_hi[3] = {-1, -1, -1};  // padded root frame (inverted = nothing to hit)
  bool built = false;
};

// threads <= 0: hardware concurrency
void build_mesh(MeshBuild &m, int threads = 0);

}  // namespace pt
