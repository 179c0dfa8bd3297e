"""Autograd binding of the B200 soft rasterizer -- the replacement for the reference's `GenDRFunction`
(/root/reference/gendr/functional/renderer.py:11-236) and `render()` (:239-288).

Same call surface (names, defaults, name-or-id arguments, asserts), but the lean path: outputs come from
torch.empty and are fully written by one forward launch (the reference does 2 clones, 3 fills and 3 in-place scales
first, renderer.py:130-151), the per-face records computed in forward are kept for backward, and texture gradients
are only produced when `textures.requires_grad`.
"""
import torch
from torch.autograd import Function

from ..cuda import generalized_renderer as _ext

# name -> id maps of the reference (functional/renderer.py:44-83)
DIST_FUNC_IDS = {
    'hard': 0, 'heaviside': 0,
    'uniform': 1, 'cubic_hermite': 2, 'wigner_semicircle': 3,
    'gaussian': 4, 'laplace': 5, 'logistic': 6, 'gudermannian': 7, 'hyperbolic_secant': 7,
    'cauchy': 8, 'reciprocal': 9,
    'gumbel_max': 10, 'gumbel_min': 11, 'exponential': 12, 'exponential_rev': 13,
    'gamma': 14, 'gamma_rev': 15, 'levy': 16, 'levy_rev': 17,
}
AGGR_ALPHA_FUNC_IDS = {
    'hard': 0, 'max': 1, 'probabilistic': 2, 'einstein': 3, 'hamacher': 4, 'frank': 5, 'yager': 6,
    'aczel_alsina': 7, 'dombi': 8, 'schweizer_sklar': 9,
}
AGGR_RGB_FUNC_IDS = {'hard': 0, 'softmax': 1}
TEXTURE_TYPE_IDS = {'surface': 0, 'vertex': 1}


def _resolve(value, table):
    # ints pass through (experiments/opt_shape.py:150,156 call with ids); unknown names raise KeyError like the reference
    return value if isinstance(value, int) else table[value]


class GenDRFunction(Function):
    @staticmethod
    def forward(ctx, face_vertices, textures, image_size=256, background_color=[0, 0, 0],
                dist_func='uniform', dist_scale=1e-2, dist_squared=False, dist_shape=None, dist_shift=None,
                dist_eps=1e4, aggr_alpha_func='probabilistic', aggr_alpha_t_conorm_p=None,
                aggr_rgb_func='softmax', aggr_rgb_eps=1e-3, aggr_rgb_gamma=1e-3, near=1, far=100,
                double_side=True, texture_type='surface', anti_aliasing=False):
        assert dist_scale >= 0, dist_scale      # functional/renderer.py:96
        assert dist_eps >= 1, dist_eps          # functional/renderer.py:101
        if not face_vertices.is_cuda:
            raise TypeError('GenDR only supports CUDA Tensors.')
        params = _ext.make_params(
            image_size, _resolve(dist_func, DIST_FUNC_IDS), dist_scale, dist_squared, dist_shape, dist_shift, dist_eps,
            _resolve(aggr_alpha_func, AGGR_ALPHA_FUNC_IDS), aggr_alpha_t_conorm_p,
            _resolve(aggr_rgb_func, AGGR_RGB_FUNC_IDS), aggr_rgb_eps, aggr_rgb_gamma, near, far, double_side,
            TEXTURE_TYPE_IDS[texture_type], background_color)

        faces = face_vertices.detach().to(torch.float32).contiguous()
        B, F = faces.shape[:2]
        faces = faces.view(B, F, 9)
        tex = textures.detach().to(device=faces.device, dtype=torch.float32).contiguous()
        tex = tex.view(B, F, -1, 3) if tex.numel() else tex.new_zeros((B, F, 1, 3))
        S = int(image_size)
        soft_colors = torch.empty((B, 4, S, S), dtype=torch.float32, device=faces.device)
        aggrs_info = torch.empty((B, 2, S, S), dtype=torch.float32, device=faces.device)
        workspace = _ext.workspace_for(faces)
        pooled = torch.empty((B, 4, S // 2, S // 2), dtype=torch.float32, device=faces.device) if anti_aliasing else None
        with torch.cuda.device(faces.device):
            if anti_aliasing:      # fused F.avg_pool2d(images, 2, 2) (gendr/renderer.py:92-93)
                _ext.forward_render_aa_raw(faces, tex, aggrs_info, soft_colors, pooled, params, workspace)
            else:
                _ext.forward_render_raw(faces, tex, None, aggrs_info, soft_colors, params, False, workspace)
        ctx.params, ctx.anti_aliasing = params, bool(anti_aliasing)
        ctx.shapes = (face_vertices.shape, textures.shape)
        ctx.save_for_backward(faces, tex, soft_colors, aggrs_info, workspace)
        if anti_aliasing:
            return pooled
        return soft_colors

    @staticmethod
    def backward(ctx, grad_soft_colors):
        faces, tex, soft_colors, aggrs_info, workspace = ctx.saved_tensors
        grad_soft_colors = grad_soft_colors.to(torch.float32).contiguous()
        want_tex = ctx.needs_input_grad[1]
        grad_faces = torch.empty_like(faces)
        grad_tex = torch.empty_like(tex) if want_tex else None
        with torch.cuda.device(faces.device):
            if ctx.anti_aliasing:
                _ext.backward_render_aa_raw(faces, tex, soft_colors, aggrs_info, grad_faces, grad_tex, grad_soft_colors,
                                            ctx.params, workspace, True, True)
            else:
                _ext.backward_render_raw(faces, tex, soft_colors, aggrs_info, grad_faces, grad_tex, grad_soft_colors,
                                         ctx.params, workspace, True, True)
        fshape, tshape = ctx.shapes
        return (grad_faces.view(fshape), grad_tex.view(tshape) if want_tex else None) + (None,) * 18


# Face indices are checked ONCE per index tensor object (one min/max reduction + host sync; a mesh keeps its faces tensor across
# iterations, and the mark is invalidated by in-place modification through the tensor's version counter): torch's own gather --
# vertices.reshape(B*V, 3)[faces] in gendr/functional/face_vertices.py:27 -- raises on an out-of-range index, the fused kernels
# clamp, so a malformed mesh must fail here instead of rendering silently.
def check_face_indices(faces, num_vertices):
    mark = (faces._version, int(num_vertices))
    if getattr(faces, '_gendr_validated', None) == mark:
        return
    if faces.numel():
        lo, hi = int(faces.min()), int(faces.max())
        if lo < 0 or hi >= num_vertices:
            raise IndexError('face index out of range: faces span [%d, %d] but the mesh has %d vertices' % (lo, hi, num_vertices))
    try:
        faces._gendr_validated = mark
    except AttributeError:      # exotic tensor subclasses without a __dict__: validate every call
        pass


class GenDRIndexedFunction(Function):
    """render() for an indexed mesh: (vertices [B,V,3] screen space, faces [B,F,3] or [F,3] int) instead of the gathered
    face_vertices [B,F,3,3].  The gather runs inside the face preprocessing kernel and the gradient is scatter-added
    into grad_vertices [B,V,3] by the backward kernel (replaces gendr/functional/face_vertices.py:27 and its
    index_put backward)."""
    @staticmethod
    def forward(ctx, vertices, faces, textures, image_size, background_color, dist_func, dist_scale, dist_squared, dist_shape,
                dist_shift, dist_eps, aggr_alpha_func, aggr_alpha_t_conorm_p, aggr_rgb_func, aggr_rgb_eps, aggr_rgb_gamma,
                near, far, double_side, texture_type, anti_aliasing=False):
        assert dist_scale >= 0, dist_scale
        assert dist_eps >= 1, dist_eps
        if not vertices.is_cuda:
            raise TypeError('GenDR only supports CUDA Tensors.')
        params = _ext.make_params(
            image_size, _resolve(dist_func, DIST_FUNC_IDS), dist_scale, dist_squared, dist_shape, dist_shift, dist_eps,
            _resolve(aggr_alpha_func, AGGR_ALPHA_FUNC_IDS), aggr_alpha_t_conorm_p,
            _resolve(aggr_rgb_func, AGGR_RGB_FUNC_IDS), aggr_rgb_eps, aggr_rgb_gamma, near, far, double_side,
            TEXTURE_TYPE_IDS[texture_type], background_color)
        verts = vertices.detach().to(torch.float32).contiguous()
        B, V = verts.shape[:2]
        check_face_indices(faces, V)
        index = faces.detach().to(device=verts.device, dtype=torch.int32).contiguous()
        shared = index.ndimension() == 2
        F = index.shape[-2]
        tex = textures.detach().to(device=verts.device, dtype=torch.float32).contiguous()
        tex = tex.view(B, F, -1, 3) if tex.numel() else tex.new_zeros((B, F, 1, 3))
        S = int(image_size)
        soft_colors = torch.empty((B, 4, S, S), dtype=torch.float32, device=verts.device)
        aggrs_info = torch.empty((B, 2, S, S), dtype=torch.float32, device=verts.device)
        lib = _ext._lib.load()
        workspace = torch.empty(lib.gendr_workspace_bytes(B, F), dtype=torch.uint8, device=verts.device)
        pooled = torch.empty((B, 4, S // 2, S // 2), dtype=torch.float32, device=verts.device) if anti_aliasing else None
        with torch.cuda.device(verts.device):
            _ext._lib.check(lib.gendr_forward_render_indexed(
                verts.data_ptr(), index.data_ptr(), int(shared), tex.data_ptr(), aggrs_info.data_ptr(), soft_colors.data_ptr(),
                pooled.data_ptr() if anti_aliasing else None, B, V, F, int(tex.shape[2]), params, workspace.data_ptr(),
                workspace.numel(), torch.cuda.current_stream(verts.device).cuda_stream))
        ctx.params, ctx.dims, ctx.shapes = params, (B, V, F, int(tex.shape[2]), shared), (vertices.shape, textures.shape)
        ctx.anti_aliasing = bool(anti_aliasing)
        ctx.save_for_backward(index, tex, soft_colors, aggrs_info, workspace)
        return pooled if anti_aliasing else soft_colors

    @staticmethod
    def backward(ctx, grad_soft_colors):
        index, tex, soft_colors, aggrs_info, workspace = ctx.saved_tensors
        B, V, F, T, shared = ctx.dims
        grad_soft_colors = grad_soft_colors.to(torch.float32).contiguous()
        want_tex = ctx.needs_input_grad[2]
        grad_vertices = torch.empty((B, V, 3), dtype=torch.float32, device=tex.device)
        grad_tex = torch.empty_like(tex) if want_tex else None
        lib = _ext._lib.load()
        with torch.cuda.device(tex.device):
            _ext._lib.check(lib.gendr_backward_render_indexed(
                index.data_ptr(), int(shared), tex.data_ptr(), soft_colors.data_ptr(), aggrs_info.data_ptr(),
                grad_vertices.data_ptr(), grad_tex.data_ptr() if want_tex else None, grad_soft_colors.data_ptr(),
                int(ctx.anti_aliasing), B, V, F, T, ctx.params, 1, workspace.data_ptr(), workspace.numel(),
                torch.cuda.current_stream(tex.device).cuda_stream))
        vshape, tshape = ctx.shapes
        return (grad_vertices.view(vshape), None, grad_tex.view(tshape) if want_tex else None) + (None,) * 18


def render_indexed(vertices, faces, textures, image_size=256, background_color=[0, 0, 0],
                   dist_func='uniform', dist_scale=1e-2, dist_squared=False, dist_shape=None, dist_shift=None, dist_eps=1e4,
                   aggr_alpha_func='probabilistic', aggr_alpha_t_conorm_p=None,
                   aggr_rgb_func='softmax', aggr_rgb_eps=1e-3, aggr_rgb_gamma=1e-3,
                   near=1, far=100, double_side=True, texture_type='surface', anti_aliasing=False):
    """Same as render(), for (vertices [B,V,3], faces [B,F,3] | [F,3]) instead of face_vertices [B,F,3,3]."""
    return GenDRIndexedFunction.apply(vertices, faces, textures, image_size, background_color, dist_func, dist_scale,
                                      dist_squared, dist_shape, dist_shift, dist_eps, aggr_alpha_func, aggr_alpha_t_conorm_p,
                                      aggr_rgb_func, aggr_rgb_eps, aggr_rgb_gamma, near, far, double_side, texture_type,
                                      anti_aliasing)


def render(face_vertices, textures, image_size=256, background_color=[0, 0, 0],
           dist_func='uniform', dist_scale=1e-2, dist_squared=False, dist_shape=None, dist_shift=None, dist_eps=1e4,
           aggr_alpha_func='probabilistic', aggr_alpha_t_conorm_p=None,
           aggr_rgb_func='softmax', aggr_rgb_eps=1e-3, aggr_rgb_gamma=1e-3,
           near=1, far=100, double_side=True, texture_type='surface', anti_aliasing=False):
    """face_vertices [B,F,3,3] (screen space), textures [B,F,T,3] -> RGBA images [B,4,S,S].
    Keyword surface and defaults of gendr.functional.render (functional/renderer.py:239-262).  One addition:
    anti_aliasing=True treats image_size as the supersampled side and returns the 2x2-averaged image
    [B,4,S/2,S/2] -- F.avg_pool2d(render(...), 2, 2) of gendr/renderer.py:92-93, fused into the kernels (bit-identical)."""
    return GenDRFunction.apply(face_vertices, textures, image_size, background_color, dist_func, dist_scale,
                               dist_squared, dist_shape, dist_shift, dist_eps, aggr_alpha_func, aggr_alpha_t_conorm_p,
                               aggr_rgb_func, aggr_rgb_eps, aggr_rgb_gamma, near, far, double_side, texture_type,
                               anti_aliasing)


def make_camera_params(mode='look_at', perspective=True, viewing_angle=30., viewing_scale=1.0, at=(0, 0, 0), up=(0, 1, 0),
                       direction=(0, 0, 1)):
    c = _ext._lib.CameraParams()
    c.mode = {'look_at': 0, 'look': 1}[mode]
    c.perspective = int(bool(perspective))
    c.viewing_angle, c.viewing_scale = float(viewing_angle), float(viewing_scale)
    target = at if c.mode == 0 else direction
    for k in range(3):
        c.at_or_direction[k] = float(target[k])
        c.up[k] = float(up[k])
    return c


def make_light_params(intensity_ambient=0.5, color_ambient=(1, 1, 1), intensity_directional=0.5,
                      color_directional=(1, 1, 1), direction=(0, 1, 0)):
    p = _ext._lib.LightParams()
    p.intensity_ambient, p.intensity_directional = float(intensity_ambient), float(intensity_directional)
    for k in range(3):
        p.color_ambient[k] = float(color_ambient[k])
        p.color_directional[k] = float(color_directional[k])
        p.direction[k] = float(direction[k])
    return p


class GenDRSceneFunction(Function):
    """Lighting -> LookAt/Look -> GenDR for a world-space mesh in ONE autograd node:
    the camera transform (gendr/functional/look_at.py:11-68, transform.py:14-44), the surface lighting
    (gendr/lighting.py:48-58), the vertices[faces] gather and the rasterizer run as four launches forward and three
    backward (gendr_scene_forward / gendr_scene_backward); gradients w.r.t. the world-space vertices (camera path +
    normals' path) and the unlit textures."""
    @staticmethod
    def forward(ctx, vertices, faces, textures, eyes, camera, light, params, anti_aliasing):
        if not vertices.is_cuda:
            raise TypeError('GenDR only supports CUDA Tensors.')
        verts = vertices.detach().to(torch.float32).contiguous()
        dev = verts.device
        eyes = eyes.detach().to(device=dev, dtype=torch.float32).contiguous()
        eyes_batched = eyes.ndimension() == 2
        # vertices [V,3]: ONE mesh seen from every eye (the shared-mesh pattern of experiments/opt_shape.py:86 without
        # vertices.repeat(batch, 1, 1)); the gradient comes back batch-summed as [V,3]
        verts_shared = verts.ndimension() == 2
        if verts_shared:
            V = verts.shape[0]
            B = eyes.shape[0] if eyes_batched else int(textures.shape[0])
        else:
            B, V = verts.shape[:2]
        check_face_indices(faces, V)
        index = faces.detach().to(device=dev, dtype=torch.int32).contiguous()
        shared = index.ndimension() == 2
        F = index.shape[-2]
        tex = textures.detach().to(device=dev, dtype=torch.float32).contiguous()
        tex = tex.view(B, F, -1, 3) if tex.numel() else tex.new_zeros((B, F, 1, 3))
        T = int(tex.shape[2])
        if eyes_batched and eyes.shape[0] != B:
            raise ValueError('eyes must be [3] or [batch, 3]')
        S = int(params.image_size)
        soft_colors = torch.empty((B, 4, S, S), dtype=torch.float32, device=dev)
        aggrs_info = torch.empty((B, 2, S, S), dtype=torch.float32, device=dev)
        pooled = torch.empty((B, 4, S // 2, S // 2), dtype=torch.float32, device=dev) if anti_aliasing else None
        lib = _ext._lib.load()
        workspace = torch.empty(lib.gendr_scene_workspace_bytes(B, V, F, T), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _ext._lib.check(lib.gendr_scene_forward(
                verts.data_ptr(), int(verts_shared), index.data_ptr(), int(shared), tex.data_ptr(), eyes.data_ptr(), int(eyes_batched), camera, light,
                aggrs_info.data_ptr(), soft_colors.data_ptr(), pooled.data_ptr() if anti_aliasing else None, B, V, F, T, params,
                workspace.data_ptr(), workspace.numel(), torch.cuda.current_stream(dev).cuda_stream))
        ctx.cfg = (camera, light, params, bool(anti_aliasing), (B, V, F, T, shared, eyes_batched, verts_shared), (vertices.shape, textures.shape))
        ctx.save_for_backward(verts, index, tex, eyes, soft_colors, aggrs_info, workspace)
        return pooled if anti_aliasing else soft_colors

    @staticmethod
    def backward(ctx, grad_images):
        verts, index, tex, eyes, soft_colors, aggrs_info, workspace = ctx.saved_tensors
        camera, light, params, aa, (B, V, F, T, shared, eyes_batched, verts_shared), (vshape, tshape) = ctx.cfg
        grad_images = grad_images.to(torch.float32).contiguous()
        want_tex = ctx.needs_input_grad[2]
        grad_vertices = torch.empty_like(verts)
        grad_tex = torch.empty_like(tex) if want_tex else None
        lib = _ext._lib.load()
        with torch.cuda.device(verts.device):
            _ext._lib.check(lib.gendr_scene_backward(
                verts.data_ptr(), int(verts_shared), index.data_ptr(), int(shared), tex.data_ptr(), eyes.data_ptr(), int(eyes_batched), camera, light,
                soft_colors.data_ptr(), aggrs_info.data_ptr(), grad_images.data_ptr(), int(aa), grad_vertices.data_ptr(),
                grad_tex.data_ptr() if want_tex else None, B, V, F, T, params, workspace.data_ptr(), workspace.numel(),
                torch.cuda.current_stream(verts.device).cuda_stream))
        return grad_vertices.view(vshape), None, grad_tex.view(tshape) if want_tex else None, None, None, None, None, None


def render_scene(vertices, faces, textures, eyes, camera=None, lighting=None, image_size=256, background_color=[0, 0, 0],
                 dist_func='uniform', dist_scale=1e-2, dist_squared=False, dist_shape=None, dist_shift=None, dist_eps=1e4,
                 aggr_alpha_func='probabilistic', aggr_alpha_t_conorm_p=None,
                 aggr_rgb_func='softmax', aggr_rgb_eps=1e-3, aggr_rgb_gamma=1e-3,
                 near=1, far=100, double_side=True, texture_type='surface', anti_aliasing=False):
    """World-space mesh (vertices [B,V,3], faces [B,F,3] | [F,3], surface textures [B,F,T,3]) seen from `eyes` ([B,3] | [3])
    -> RGBA images: lighting(mesh); transform(mesh); renderer(mesh) of the reference's scripts in one fused node.
    vertices [V,3] (2-D): ONE mesh shared by all views (experiments/opt_shape.py:86 renders vertices.repeat(batch, 1, 1)); the
    copies are never materialised and the vertex gradient arrives batch-summed, [V,3] -- the buffer a data-parallel rank hands
    to its single all-reduce (gendr_b200.parallel.allreduce_shared_vertex_grads).
    camera: dict for make_camera_params (mode, perspective, viewing_angle, viewing_scale, at, up, direction);
    lighting: dict for make_light_params, or None for no lighting step."""
    assert dist_scale >= 0, dist_scale
    assert dist_eps >= 1, dist_eps
    if texture_type != 'surface':
        raise ValueError('render_scene supports surface textures only')
    params = _ext.make_params(
        image_size, _resolve(dist_func, DIST_FUNC_IDS), dist_scale, dist_squared, dist_shape, dist_shift, dist_eps,
        _resolve(aggr_alpha_func, AGGR_ALPHA_FUNC_IDS), aggr_alpha_t_conorm_p,
        _resolve(aggr_rgb_func, AGGR_RGB_FUNC_IDS), aggr_rgb_eps, aggr_rgb_gamma, near, far, double_side,
        TEXTURE_TYPE_IDS[texture_type], background_color)
    cam = make_camera_params(**(camera or {}))
    light = make_light_params(**lighting) if lighting is not None else None
    if not torch.is_tensor(eyes):
        eyes = torch.tensor(eyes, dtype=torch.float32, device=vertices.device)
    if eyes.requires_grad:
        raise ValueError('render_scene does not differentiate w.r.t. the camera position: use LookAt/Look (torch path) for eyes that '
                         'require a gradient (experiments/opt_camera.py)')
    return GenDRSceneFunction.apply(vertices, faces, textures, eyes, cam, light, params, anti_aliasing)
