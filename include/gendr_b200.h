/* gendr_b200.h -- C ABI of the B200-native GenDR soft rasterizer (libgendr_b200.so).
 *
 * Drop-in boundary.  These entry points are what the reference's Python binding for the hot path binds today via
 * pybind11 (module `gendr.cuda.generalized_renderer`, /root/reference/gendr/cuda/generalized_renderer_cuda.cpp):
 *
 *   gendr_forward_render    replaces  forward_render   (generalized_renderer_cuda.cpp:74-127  -> K.cu:1071-1152)
 *   gendr_backward_render   replaces  backward_render  (generalized_renderer_cuda.cpp:130-192 -> K.cu:1155-1227)
 *   gendr_sigmoid_forward / gendr_sigmoid_backward / gendr_t_conorm_forward / gendr_t_conorm_backward
 *                           replace   sigmoid_forward ... t_conorm_backward (generalized_renderer_cuda.cpp:195-236)
 *
 * Plain pointers and sizes only; no torch types.  All device pointers are fp32, contiguous, on ONE device (the
 * library switches to the device that owns `faces`); work is enqueued on `stream` (a cudaStream_t passed as
 * void*, NULL = legacy default stream) and nothing synchronises.  Return value: 0 on success, otherwise a
 * cudaError_t (or GENDR_ERR_*); gendr_last_error() gives the text.  Unlike the reference (which only printf()s
 * launch errors, K.cu:1111-1113) errors are reported to the caller.
 *
 * Buffer conventions are the reference's (gendr/functional/renderer.py:130-151, :191-197):
 *   faces        [B,F,3,3]  screen-space (x,y in NDC [-1,1], z = depth) vertices of every face
 *   textures     [B,F,T,3]  T = texture_res^2 texels (surface) or T = 3 vertex colours (vertex)
 *   soft_colors  [B,4,S,S]  planar RGBA output
 *   aggrs_info   [B,2,S,S]  softmax (sum,max) or hard (depth_min, face_index_min)
 *   faces_info   [B,F,27]   optional: the reference's per-face scratch (inv 9 | gram 9 | obtuse 3 | 6 unused)
 *   grad_faces   [B,F,3,3], grad_textures [B,F,T,3], grad_soft_colors [B,4,S,S]
 */
#ifndef GENDR_B200_H
#define GENDR_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GENDR_ERR_INVALID_ARGUMENT 100001
#define GENDR_ERR_WORKSPACE_TOO_SMALL 100002

/* The 16 scalars of forward_render/backward_render (generalized_renderer_cuda.cpp:80-95), plus the background
 * colour that the reference passes pre-filled inside soft_colors (functional/renderer.py:149-151). */
typedef struct gendr_render_params {
    int   image_size;
    int   dist_func;              /* 0..17, functional/renderer.py:44-63 */
    float dist_scale;
    int   dist_squared;
    float dist_shape;
    float dist_shift;
    float dist_eps;
    int   aggr_alpha_func;        /* 0..9, functional/renderer.py:68-79 */
    float aggr_alpha_t_conorm_p;
    int   aggr_rgb_func;          /* 0 hard, 1 softmax */
    float aggr_rgb_eps;
    float aggr_rgb_gamma;
    float near_plane;
    float far_plane;
    int   double_side;
    int   texture_type;           /* 0 surface, 1 vertex */
    float background[3];
} gendr_render_params;

/* Scratch the library needs between forward and backward: per-face records (176 B) + packed pixel rects (8 B), and per batch
 * item 32 KB for the tile counters and the longest-first CTA order of large grids (up to 4096 tiles per image). */
size_t gendr_workspace_bytes(int batch, int faces);

/* Forward.  background_prefilled != 0: read the background from soft_colors' RGB planes exactly as the reference
 * does; 0: use params->background and treat soft_colors as uninitialised output.  faces_info may be NULL. */
int gendr_forward_render(const float* faces, const float* textures, float* faces_info, float* aggrs_info,
                         float* soft_colors, int batch, int num_faces, int texture_size,
                         const gendr_render_params* params, int background_prefilled,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Backward.  workspace_valid != 0: the workspace still holds what gendr_forward_render left there for the same
 * faces/params (skips re-running the face preprocessing).  zero_grads != 0: the library zero-fills grad_faces /
 * grad_textures first (the reference expects the caller to pass zeros).  grad_textures may be NULL. */
int gendr_backward_render(const float* faces, const float* textures, const float* soft_colors,
                          const float* aggrs_info, float* grad_faces, float* grad_textures,
                          const float* grad_soft_colors, int batch, int num_faces, int texture_size,
                          const gendr_render_params* params, int workspace_valid, int zero_grads,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Batch-summed gradient (SURVEY.md 8(e) "fusion with the collective").  When one mesh is shared by the whole batch
 * (vertices.repeat(batch, 1, 1), /root/reference/experiments/opt_shape.py:86) the gradient the optimiser needs is the SUM over the
 * batch -- which the reference gets from autograd's backward of `repeat` after writing [B,F,3,3].  These variants accumulate
 * straight into ONE [F,3,3] (or, indexed, [V,3]) buffer, so a data-parallel rank can hand that buffer to a single all-reduce
 * without an intermediate [B,F,3,3] tensor or a reduction kernel.  Everything else as in gendr_backward_render(_indexed). */
int gendr_backward_render_batchsum(const float* faces, const float* textures, const float* soft_colors,
                                   const float* aggrs_info, float* grad_faces_sum, float* grad_textures,
                                   const float* grad_soft_colors, int batch, int num_faces, int texture_size,
                                   const gendr_render_params* params, int workspace_valid, int zero_grads,
                                   void* workspace, size_t workspace_bytes, void* stream);
int gendr_backward_render_indexed_batchsum(const int* face_index, int index_shared, const float* textures, const float* soft_colors,
                                           const float* aggrs_info, float* grad_vertices_sum, float* grad_textures,
                                           const float* grad_soft_colors, int grad_is_pooled, int batch, int num_vertices,
                                           int num_faces, int texture_size, const gendr_render_params* params, int zero_grads,
                                           void* workspace, size_t workspace_bytes, void* stream);

/* Indexed-mesh variants: fuse the reference's `vertices[faces]` gather
 * (gendr/functional/face_vertices.py:9-27, called from gendr/mesh.py:102) into the face preprocessing and its backward
 * (a scatter-add into the vertex gradient) into the backward kernel, so the [B,F,3,3] face-vertex tensor and its
 * gradient are never materialised.  vertices [B,V,3] screen space; face_index int32 [B,F,3], or [F,3] when
 * index_shared != 0; indices are clamped to [0, V-1].  The backward must follow the forward on the same workspace.
 * pooled_colors (may be NULL) / grad_is_pooled: fused 2x anti-aliasing, see gendr_forward_render_aa. */
int gendr_forward_render_indexed(const float* vertices, const int* face_index, int index_shared, const float* textures,
                                 float* aggrs_info, float* soft_colors, float* pooled_colors, int batch, int num_vertices,
                                 int num_faces, int texture_size, const gendr_render_params* params, void* workspace,
                                 size_t workspace_bytes, void* stream);
int gendr_backward_render_indexed(const int* face_index, int index_shared, const float* textures, const float* soft_colors,
                                  const float* aggrs_info, float* grad_vertices, float* grad_textures,
                                  const float* grad_soft_colors, int grad_is_pooled, int batch, int num_vertices,
                                  int num_faces, int texture_size, const gendr_render_params* params, int zero_grads,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* Fused 2x anti-aliasing.  The reference renders at 2S and then calls
 * F.avg_pool2d(images, kernel_size=2, stride=2) (gendr/renderer.py:68,92-93,96,121-122).  Here params->image_size is the
 * supersampled side (even); the forward kernel's epilogue additionally writes pooled_colors [B,4,S/2,S/2] -- summed in
 * torch's order, so bit-identical to the unfused pooling -- and the backward kernel reads the cotangent of the POOLED
 * image, grad_pooled_colors [B,4,S/2,S/2] (avg_pool2d's backward, grad/4 per pixel, happens in its prologue).
 * soft_colors [B,4,S,S] (full resolution) is still written: the backward pass needs it. */
int gendr_forward_render_aa(const float* faces, const float* textures, float* aggrs_info, float* soft_colors,
                            float* pooled_colors, int batch, int num_faces, int texture_size,
                            const gendr_render_params* params, void* workspace, size_t workspace_bytes, void* stream);
int gendr_backward_render_aa(const float* faces, const float* textures, const float* soft_colors, const float* aggrs_info,
                             float* grad_faces, float* grad_textures, const float* grad_pooled_colors, int batch,
                             int num_faces, int texture_size, const gendr_render_params* params, int workspace_valid,
                             int zero_grads, void* workspace, size_t workspace_bytes, void* stream);

/* Camera transform and lighting as CUDA kernels, with their backward passes.
 *   gendr_camera_*    replace  gendr.functional.look_at / look (gendr/functional/look_at.py:11-68, look.py:11-56) followed by
 *                     perspective / orthogonal (gendr/transform.py:14-44), i.e. LookAt.transform / Look.transform
 *                     (transform.py:132-138, :161-168) and what autograd derives for them
 *   gendr_lighting_*  replace  gendr.Lighting.forward for surface textures (gendr/lighting.py:48-58: ambient + one
 *                     directional light from the face normal, gendr/mesh.py:104-108) and its autograd backward
 * eyes: device [B,3] (eyes_batched != 0) or [3].  All buffers fp32 device memory except face_index (int32). */
typedef struct gendr_camera_params {
    int   mode;               /* 0 = look_at (camera looks from eye to at_or_direction), 1 = look (along at_or_direction) */
    int   perspective;        /* 1 = perspective(viewing_angle), 0 = orthogonal(viewing_scale) */
    float viewing_angle;      /* degrees, transform.py:14 */
    float viewing_scale;      /* transform.py:32 */
    float at_or_direction[3];
    float up[3];
} gendr_camera_params;
typedef struct gendr_light_params {
    float intensity_ambient;      float color_ambient[3];       /* lighting.py:38 */
    float intensity_directional;  float color_directional[3];   /* lighting.py:39 */
    float direction[3];                                          /* lighting.py:40 */
} gendr_light_params;

int gendr_camera_forward(const float* vertices, const float* eyes, int eyes_batched, float* screen_vertices, int batch,
                         int num_vertices, const gendr_camera_params* camera, void* stream);
/* grad_vertices [B,V,3] is overwritten (no zero-fill needed).  grad_eyes (may be NULL): gradient w.r.t. the camera position -- the
 * optimisation target of /root/reference/experiments/opt_camera.py (eye with requires_grad, :236), which the reference gets from
 * autograd through look_at() -- [B,3], or [3] summed over the batch when one eye is shared; eye_scratch: [B,12] floats of device
 * scratch, required with grad_eyes (per-item accumulators, reduced by a second tiny kernel). */
int gendr_camera_backward(const float* vertices, const float* eyes, int eyes_batched, const float* grad_screen_vertices,
                          float* grad_vertices, float* grad_eyes, float* eye_scratch, int batch, int num_vertices,
                          const gendr_camera_params* camera, void* stream);
int gendr_lighting_forward(const float* vertices, const int* face_index, int index_shared, const float* textures,
                           float* lit_textures, int batch, int num_vertices, int num_faces, int texture_size,
                           const gendr_light_params* light, void* stream);
/* grad_textures [B,F,T,3] is overwritten (may be NULL); the normal's gradient is ADDED to grad_vertices (may be NULL) */
int gendr_lighting_backward(const float* vertices, const int* face_index, int index_shared, const float* textures,
                            const float* grad_lit_textures, float* grad_textures, float* grad_vertices, int batch,
                            int num_vertices, int num_faces, int texture_size, const gendr_light_params* light, void* stream);

/* Lighting of VERTEX textures: replaces gendr.Lighting.forward for texture_type == 'vertex' (gendr/lighting.py:60-66) together with
 * the vertex normals it reads (Mesh.vertex_normals = gendr/functional/vertex_normals.py:11-49: per-corner cross products summed per
 * vertex by index_add_, then F.normalize) and what autograd derives for them.  textures / lit_textures [B,V,3].
 * normal_sums [B,V,3]: written by the forward call (zero-filled inside), read by the backward call.
 * Backward: grad_textures [B,V,3] is overwritten (may be NULL); the normals' gradient is ADDED to grad_vertices [B,V,3] (may be
 * NULL; otherwise grad_sums_scratch [B,V,3] floats of device scratch is required). */
int gendr_vertex_lighting_forward(const float* vertices, const int* face_index, int index_shared, const float* textures,
                                  float* lit_textures, float* normal_sums, int batch, int num_vertices, int num_faces,
                                  const gendr_light_params* light, void* stream);
int gendr_vertex_lighting_backward(const float* vertices, const int* face_index, int index_shared, const float* textures,
                                   const float* normal_sums, const float* grad_lit_textures, float* grad_textures,
                                   float* grad_vertices, float* grad_sums_scratch, int batch, int num_vertices, int num_faces,
                                   const gendr_light_params* light, void* stream);

/* The whole scene step in one call: world-space mesh -> Lighting -> LookAt/Look -> GenDR (the order every script of the
 * reference uses: experiments/opt_shape.py:257-259, train_reconstruction.py:228-230), surface textures.
 * Forward: camera kernel, lighting kernel, indexed face preprocessing, render kernel (4 launches; the reference issues ~40).
 * Backward: render backward (scatter-adds into the screen-space vertex gradient and the lit-texture gradient held in the
 * workspace), camera backward, lighting backward -> grad_vertices [B,V,3] w.r.t. the WORLD-space vertices and
 * grad_textures [B,F,T,3] w.r.t. the UNLIT textures (may be NULL), grad_eyes w.r.t. the camera position(s) (may be NULL; see
 * gendr_camera_backward).  light may be NULL (no lighting step).
 * pooled_colors / grad_is_pooled as in gendr_forward_render_aa.
 * vertices_shared != 0: ONE world-space mesh `vertices` [V,3] seen from `batch` eyes (the shared-mesh pattern of
 * experiments/opt_shape.py:86 without materialising vertices.repeat(batch,1,1)); gendr_scene_backward then returns the
 * batch-SUMMED gradient grad_vertices [V,3] (camera path: red.add per view; normals' path: added once per view), which is what
 * autograd's backward of `repeat` would produce -- ready for one all-reduce across data-parallel ranks. */
size_t gendr_scene_workspace_bytes(int batch, int num_vertices, int num_faces, int texture_size);
int gendr_scene_forward(const float* vertices, int vertices_shared, const int* face_index, int index_shared, const float* textures,
                        const float* eyes, int eyes_batched, const gendr_camera_params* camera,
                        const gendr_light_params* light, float* aggrs_info, float* soft_colors, float* pooled_colors,
                        int batch, int num_vertices, int num_faces, int texture_size, const gendr_render_params* params,
                        void* workspace, size_t workspace_bytes, void* stream);
int gendr_scene_backward(const float* vertices, int vertices_shared, const int* face_index, int index_shared, const float* textures,
                         const float* eyes, int eyes_batched, const gendr_camera_params* camera,
                         const gendr_light_params* light, const float* soft_colors, const float* aggrs_info,
                         const float* grad_soft_colors, int grad_is_pooled, float* grad_vertices, float* grad_textures,
                         float* grad_eyes, int batch, int num_vertices, int num_faces, int texture_size,
                         const gendr_render_params* params, void* workspace, size_t workspace_bytes, void* stream);

/* Voxelizer: replaces gendr.functional.voxelization(faces, size, normalize=False)
 * (gendr/functional/voxelization.py:45-62) -- the pybind functions voxelize_sub1..4 of gendr/cuda/voxelization_cuda.cpp:85-88,
 * the three axis permutations, the threshold and the host loop with two .sum() synchronisations per flood-fill sweep -- by
 * two launches without any host synchronisation.  faces [B,F,3,3] fp32 in unit-cube coordinates (what Mesh.voxelize passes,
 * gendr/mesh.py:124-126; scaled by voxel_size inside), voxels int32 [B,vs,vs,vs] (1 = surface or inside, 0 = outside),
 * fully overwritten.  Bit-identical to the reference's CUDA kernels. */
size_t gendr_voxelize_workspace_bytes(int batch, int voxel_size);
int gendr_voxelize(const float* faces, int* voxels, int batch, int num_faces, int voxel_size, void* workspace,
                   size_t workspace_bytes, void* stream);

/* End-to-end convenience with HOST buffers (pinned or pageable): H2D copies of faces/textures/grad_soft_colors,
 * forward + backward on the current device, D2H copies of soft_colors/grad_faces/grad_textures, one stream
 * synchronisation at the end.  Device scratch is cached inside the library between calls (per device). */
int gendr_render_forward_backward_host(const float* h_faces, const float* h_textures, const float* h_grad_soft_colors,
                                       float* h_soft_colors, float* h_grad_faces, float* h_grad_textures,
                                       int batch, int num_faces, int texture_size, const gendr_render_params* params);
/* The scratch (device buffers, three streams, events) is kept PER DEVICE -- the call uses the current device's -- so one process
 * may drive several GPUs; this frees all of it (every device), e.g. before a fork or at shutdown. */
void gendr_release_host_scratch(void);

/* Scalar functions of the reference module.  Like the reference's (host instantiations of __host__ __device__ templates,
 * K.cu:1230-1270) they run on the CPU: no device, no launch, no allocation -- animations/distributions_to_csv.py:19 calls them
 * thousands of times.  They are the host instantiation of the same templates the render kernels use. */
float gendr_sigmoid_forward(int function_id, float sign, float x, float scale, float dist_shape, float dist_shift);
float gendr_sigmoid_backward(int function_id, float sign, float x, float scale, float dist_shape, float dist_shift);
float gendr_t_conorm_forward(int t_conorm_id, float a_existing, float b_new, int face_id, float t_conorm_p);
float gendr_t_conorm_backward(int t_conorm_id, float a_all, float b_current, int number_of_faces, float t_conorm_p);

/* Diagnostics. */
const char* gendr_last_error(void);
const char* gendr_version(void);
/* number of kernels this library has launched since load (bench.py's "gpu_launches") */
long long gendr_launch_count(void);
/* self-test: the device instantiation of a scalar function, evaluated by ONE GPU thread (what: 0 sigmoid_forward, 1 sigmoid_backward,
 * 2 t_conorm_forward, 3 t_conorm_backward; a, b = (sign, x) or (a, b)); tests compare it with the host functions above */
float gendr_selftest_scalar_device(int what, int id, float a, float b, float scale, float shape, float shift, float p);
/* self-test: number of (dividend, divisor) pairs out of n pseudo-random ones for which the library's shared-reciprocal
 * division differs from IEEE division by a single bit (must be 0); -1 on CUDA error */
long long gendr_selftest_division(long long n);
/* geometry probe: one thread evaluates prep + barycentric + projection for n (face, pixel) pairs.
 * faces [n,9], xy [n,2] -> out [n,10] = w0 w1 w2 t0 t1 t2 dx dy sign d2   (device pointers) */
int gendr_probe_pairs(const float* faces, const float* xy, float* out, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GENDR_B200_H */
