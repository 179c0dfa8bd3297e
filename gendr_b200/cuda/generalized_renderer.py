"""Drop-in for the reference extension module `gendr.cuda.generalized_renderer`
(/root/reference/gendr/cuda/generalized_renderer_cuda.cpp:230-237): same six functions, same positional
signatures, same in-place / return conventions, backed by libgendr_b200.so through its C ABI.

    forward_render(faces, textures, faces_info, aggrs_info, soft_colors, image_size, dist_func, dist_scale,
                   dist_squared, dist_shape, dist_shift, dist_eps, aggr_alpha_func, aggr_alpha_t_conorm_p,
                   aggr_rgb_func, aggr_rgb_eps, aggr_rgb_gamma, near, far, double_side, texture_type)
        -> [faces_info, aggrs_info, soft_colors]                      (generalized_renderer_cuda.cpp:74-127)
    backward_render(faces, textures, soft_colors, faces_info, aggrs_info, grad_faces, grad_textures,
                    grad_soft_colors, <same 16 scalars>) -> [grad_faces, grad_textures]      (:130-192)
    sigmoid_forward / sigmoid_backward / t_conorm_forward / t_conorm_backward                 (:195-236)

Differences, all deliberate: kernels run on the tensors' device and on torch's current stream (the reference uses
the legacy default stream of the current device, K.cu:1103,1118,1190); CUDA errors raise instead of being
printf()ed (K.cu:1111-1113); `None` for the three optional shape parameters means 0.0 (the reference's pybind
signature rejects None).
"""
import torch

from .. import _lib


def _check_input(t, name):
    # CHECK_CUDA / CHECK_CONTIGUOUS of the reference (generalized_renderer_cuda.cpp:69-71) -> RuntimeError
    if not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor' % name)
    if not t.is_contiguous():
        raise RuntimeError('%s must be contiguous' % name)
    if t.dtype != torch.float32:
        raise RuntimeError('%s must be float32 (the reference Python path only supports fp32, SURVEY Q6)' % name)


def _f(v):
    return 0.0 if v is None else float(v)


def make_params(image_size, dist_func, dist_scale, dist_squared, dist_shape, dist_shift, dist_eps, aggr_alpha_func,
                aggr_alpha_t_conorm_p, aggr_rgb_func, aggr_rgb_eps, aggr_rgb_gamma, near, far, double_side,
                texture_type, background=(0.0, 0.0, 0.0)):
    p = _lib.RenderParams()
    p.image_size = int(image_size)
    p.dist_func = int(dist_func); p.dist_scale = float(dist_scale); p.dist_squared = int(bool(dist_squared))
    p.dist_shape = _f(dist_shape); p.dist_shift = _f(dist_shift); p.dist_eps = float(dist_eps)
    p.aggr_alpha_func = int(aggr_alpha_func); p.aggr_alpha_t_conorm_p = _f(aggr_alpha_t_conorm_p)
    p.aggr_rgb_func = int(aggr_rgb_func); p.aggr_rgb_eps = float(aggr_rgb_eps); p.aggr_rgb_gamma = float(aggr_rgb_gamma)
    p.near_plane = float(near); p.far_plane = float(far); p.double_side = int(bool(double_side))
    p.texture_type = int(texture_type)
    p.background[0], p.background[1], p.background[2] = (float(background[0]), float(background[1]), float(background[2]))
    return p


def workspace_for(faces):
    """Scratch tensor the kernels keep between forward and backward (face records + packed pixel rects)."""
    lib = _lib.load()
    n = lib.gendr_workspace_bytes(int(faces.shape[0]), int(faces.shape[1]))
    return torch.empty(n, dtype=torch.uint8, device=faces.device)


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def forward_render_raw(faces, textures, faces_info, aggrs_info, soft_colors, params, background_prefilled, workspace):
    lib = _lib.load()
    B, F = int(faces.shape[0]), int(faces.shape[1])
    T = int(textures.shape[2])
    _lib.check(lib.gendr_forward_render(
        faces.data_ptr(), textures.data_ptr(), faces_info.data_ptr() if faces_info is not None else None,
        aggrs_info.data_ptr(), soft_colors.data_ptr(), B, F, T, params, int(background_prefilled),
        workspace.data_ptr(), workspace.numel(), _stream(faces.device)))


def backward_render_raw(faces, textures, soft_colors, aggrs_info, grad_faces, grad_textures, grad_soft_colors, params,
                        workspace, workspace_valid, zero_grads):
    lib = _lib.load()
    B, F = int(faces.shape[0]), int(faces.shape[1])
    T = int(textures.shape[2])
    _lib.check(lib.gendr_backward_render(
        faces.data_ptr(), textures.data_ptr(), soft_colors.data_ptr(), aggrs_info.data_ptr(), grad_faces.data_ptr(),
        grad_textures.data_ptr() if grad_textures is not None else None, grad_soft_colors.data_ptr(), B, F, T, params,
        int(workspace_valid), int(zero_grads), workspace.data_ptr(), workspace.numel(), _stream(faces.device)))


def backward_render_batchsum_raw(faces, textures, soft_colors, aggrs_info, grad_faces_sum, grad_textures, grad_soft_colors, params, workspace,
                                 workspace_valid, zero_grads):
    """Backward pass accumulating the gradient of a mesh shared by the batch straight into grad_faces_sum [F,9]
    (gendr_backward_render_batchsum; SURVEY 8(e) "fusion with the collective")."""
    lib = _lib.load()
    B, F = int(faces.shape[0]), int(faces.shape[1])
    T = int(textures.shape[2])
    _lib.check(lib.gendr_backward_render_batchsum(
        faces.data_ptr(), textures.data_ptr(), soft_colors.data_ptr(), aggrs_info.data_ptr(), grad_faces_sum.data_ptr(),
        grad_textures.data_ptr() if grad_textures is not None else None, grad_soft_colors.data_ptr(), B, F, T, params,
        int(workspace_valid), int(zero_grads), workspace.data_ptr(), workspace.numel(), _stream(faces.device)))


def forward_render_aa_raw(faces, textures, aggrs_info, soft_colors, pooled_colors, params, workspace):
    lib = _lib.load()
    B, F = int(faces.shape[0]), int(faces.shape[1])
    _lib.check(lib.gendr_forward_render_aa(
        faces.data_ptr(), textures.data_ptr(), aggrs_info.data_ptr(), soft_colors.data_ptr(), pooled_colors.data_ptr(), B, F,
        int(textures.shape[2]), params, workspace.data_ptr(), workspace.numel(), _stream(faces.device)))


def backward_render_aa_raw(faces, textures, soft_colors, aggrs_info, grad_faces, grad_textures, grad_pooled_colors, params,
                           workspace, workspace_valid, zero_grads):
    lib = _lib.load()
    B, F = int(faces.shape[0]), int(faces.shape[1])
    _lib.check(lib.gendr_backward_render_aa(
        faces.data_ptr(), textures.data_ptr(), soft_colors.data_ptr(), aggrs_info.data_ptr(), grad_faces.data_ptr(),
        grad_textures.data_ptr() if grad_textures is not None else None, grad_pooled_colors.data_ptr(), B, F,
        int(textures.shape[2]), params, int(workspace_valid), int(zero_grads), workspace.data_ptr(), workspace.numel(),
        _stream(faces.device)))


def forward_render(faces, textures, faces_info, aggrs_info, soft_colors, image_size, dist_func, dist_scale,
                   dist_squared, dist_shape, dist_shift, dist_eps, aggr_alpha_func, aggr_alpha_t_conorm_p,
                   aggr_rgb_func, aggr_rgb_eps, aggr_rgb_gamma, near, far, double_side, texture_type):
    for t, n in ((faces, 'faces'), (textures, 'textures'), (faces_info, 'faces_info'), (aggrs_info, 'aggrs_info'),
                 (soft_colors, 'soft_colors')):
        _check_input(t, n)
    params = make_params(image_size, dist_func, dist_scale, dist_squared, dist_shape, dist_shift, dist_eps,
                         aggr_alpha_func, aggr_alpha_t_conorm_p, aggr_rgb_func, aggr_rgb_eps, aggr_rgb_gamma, near, far,
                         double_side, texture_type)
    ws = workspace_for(faces)
    forward_render_raw(faces, textures, faces_info, aggrs_info, soft_colors, params, True, ws)
    return [faces_info, aggrs_info, soft_colors]


def backward_render(faces, textures, soft_colors, faces_info, aggrs_info, grad_faces, grad_textures, grad_soft_colors,
                    image_size, dist_func, dist_scale, dist_squared, dist_shape, dist_shift, dist_eps, aggr_alpha_func,
                    aggr_alpha_t_conorm_p, aggr_rgb_func, aggr_rgb_eps, aggr_rgb_gamma, near, far, double_side,
                    texture_type):
    for t, n in ((faces, 'faces'), (textures, 'textures'), (soft_colors, 'soft_colors'), (faces_info, 'faces_info'),
                 (aggrs_info, 'aggrs_info'), (grad_faces, 'grad_faces'), (grad_textures, 'grad_textures'),
                 (grad_soft_colors, 'grad_soft_colors')):
        _check_input(t, n)
    params = make_params(image_size, dist_func, dist_scale, dist_squared, dist_shape, dist_shift, dist_eps,
                         aggr_alpha_func, aggr_alpha_t_conorm_p, aggr_rgb_func, aggr_rgb_eps, aggr_rgb_gamma, near, far,
                         double_side, texture_type)
    # faces_info (the reference's per-face scratch) is accepted for signature compatibility; our kernels rebuild
    # their own face records from `faces` (one tiny launch), and accumulate into the caller's zero-filled grads.
    ws = workspace_for(faces)
    backward_render_raw(faces, textures, soft_colors, aggrs_info, grad_faces, grad_textures, grad_soft_colors, params,
                        ws, False, False)
    return [grad_faces, grad_textures]


def sigmoid_forward(function_id, sign, x, scale, dist_shape, dist_shift):
    return float(_lib.load().gendr_sigmoid_forward(int(function_id), sign, x, scale, dist_shape, dist_shift))


def sigmoid_backward(function_id, sign, x, scale, dist_shape, dist_shift):
    return float(_lib.load().gendr_sigmoid_backward(int(function_id), sign, x, scale, dist_shape, dist_shift))


def t_conorm_forward(t_conorm_id, a_existing, b_new, face_id, t_conorm_p):
    return float(_lib.load().gendr_t_conorm_forward(int(t_conorm_id), a_existing, b_new, int(face_id), t_conorm_p))


def t_conorm_backward(t_conorm_id, a_all, b_current, number_of_faces, t_conorm_p):
    return float(_lib.load().gendr_t_conorm_backward(int(t_conorm_id), a_all, b_current, int(number_of_faces), t_conorm_p))
