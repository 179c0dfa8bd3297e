// torch_binding.cpp -- the autograd node of functional.render() in C++ (module gendr_b200._torchbind).
//
// The reference binds its kernels through a pybind11 module and wraps them in a PYTHON autograd.Function
// (/root/reference/gendr/functional/renderer.py:11-236); for tiny scenes (1 triangle at 32x32) that Python round trip --
// Function.apply, ctx bookkeeping, the backward dispatch from the autograd engine thread -- costs more than the kernels
// (SURVEY section 7 "hard part 5").  This file is the same node written against torch's C++ autograd API: allocation of the
// outputs, the call into the C-ABI library (include/gendr_b200.h) and save_for_backward happen without touching the
// interpreter, and the backward pass runs on the autograd engine thread without taking the GIL.
//
// It is HOST code only: all device work stays in libgendr_b200.so behind the C ABI (this module links against it).  The
// ctypes path of gendr_b200/functional/renderer.py drives the same library and stays the binding of record; this module is
// used when it has been built (make -C gendr_b200/csrc torchbind).
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <cstring>
#include <string>

#include "../../include/gendr_b200.h"

namespace {

using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

inline at::Tensor f32c(const at::Tensor& t, const at::Device& dev) {
    if (t.scalar_type() == at::kFloat && t.device() == dev && t.is_contiguous()) return t;
    return t.to(dev, at::kFloat).contiguous();
}

struct RenderFaces : public torch::autograd::Function<RenderFaces> {
    // params: the bytes of a gendr_render_params struct (copied: the node owns its configuration)
    static at::Tensor forward(AutogradContext* ctx, const at::Tensor& face_vertices, const at::Tensor& textures, int64_t params_addr,
                              bool anti_aliasing) {
        TORCH_CHECK_TYPE(face_vertices.is_cuda(), "GenDR only supports CUDA Tensors.");
        const gendr_render_params params = *reinterpret_cast<const gendr_render_params*>(params_addr);
        const at::Device dev = face_vertices.device();
        at::Tensor faces = f32c(face_vertices, dev);
        const int64_t B = faces.size(0), F = faces.size(1);
        at::Tensor tex = f32c(textures, dev);
        if (tex.numel() == 0) tex = at::zeros({B, F, 1, 3}, faces.options());
        else if (tex.dim() != 4) tex = tex.view({B, F, -1, 3});
        const int64_t S = params.image_size, T = tex.size(2);
        c10::cuda::CUDAGuard guard(dev);
        at::Tensor colors = at::empty({B, 4, S, S}, faces.options());
        at::Tensor aggrs = at::empty({B, 2, S, S}, faces.options());
        at::Tensor ws = at::empty({(int64_t)gendr_workspace_bytes((int)B, (int)F)}, faces.options().dtype(at::kByte));
        at::Tensor pooled;
        void* stream = c10::cuda::getCurrentCUDAStream(dev.index()).stream();
        int rc;
        if (anti_aliasing) {      // fused F.avg_pool2d(images, 2, 2) (gendr/renderer.py:92-93)
            pooled = at::empty({B, 4, S / 2, S / 2}, faces.options());
            rc = gendr_forward_render_aa(faces.data_ptr<float>(), tex.data_ptr<float>(), aggrs.data_ptr<float>(), colors.data_ptr<float>(),
                                         pooled.data_ptr<float>(), (int)B, (int)F, (int)T, &params, ws.data_ptr(), (size_t)ws.numel(), stream);
        } else {
            rc = gendr_forward_render(faces.data_ptr<float>(), tex.data_ptr<float>(), nullptr, aggrs.data_ptr<float>(), colors.data_ptr<float>(),
                                      (int)B, (int)F, (int)T, &params, 0, ws.data_ptr(), (size_t)ws.numel(), stream);
        }
        TORCH_CHECK(rc == 0, gendr_last_error(), " (code ", rc, ")");
        ctx->save_for_backward({faces, tex, colors, aggrs, ws});
        ctx->saved_data["params"] = std::string(reinterpret_cast<const char*>(&params), sizeof params);
        ctx->saved_data["aa"] = anti_aliasing;
        ctx->saved_data["fshape"] = face_vertices.sizes().vec();
        ctx->saved_data["tshape"] = textures.sizes().vec();
        return anti_aliasing ? pooled : colors;
    }

    static variable_list backward(AutogradContext* ctx, variable_list grad_outputs) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor &faces = saved[0], &tex = saved[1], &colors = saved[2], &aggrs = saved[3], &ws = saved[4];
        gendr_render_params params;
        const std::string bytes = ctx->saved_data["params"].toStringRef();
        std::memcpy(&params, bytes.data(), sizeof params);
        const bool aa = ctx->saved_data["aa"].toBool();
        const auto fshape = ctx->saved_data["fshape"].toIntVector(), tshape = ctx->saved_data["tshape"].toIntVector();
        const at::Device dev = faces.device();
        const at::Tensor grad = f32c(grad_outputs[0], dev);
        const int64_t B = faces.size(0), F = faces.size(1), T = tex.size(2);
        const bool want_tex = ctx->needs_input_grad(1);
        c10::cuda::CUDAGuard guard(dev);
        at::Tensor grad_faces = at::empty(fshape, faces.options());
        at::Tensor grad_tex = want_tex ? at::empty_like(tex) : at::Tensor();
        void* stream = c10::cuda::getCurrentCUDAStream(dev.index()).stream();
        auto fn = aa ? gendr_backward_render_aa : gendr_backward_render;
        const int rc = fn(faces.data_ptr<float>(), tex.data_ptr<float>(), colors.data_ptr<float>(), aggrs.data_ptr<float>(), grad_faces.data_ptr<float>(),
                          want_tex ? grad_tex.data_ptr<float>() : nullptr, grad.data_ptr<float>(), (int)B, (int)F, (int)T, &params, 1, 1, ws.data_ptr(),
                          (size_t)ws.numel(), stream);
        TORCH_CHECK(rc == 0, gendr_last_error(), " (code ", rc, ")");
        if (want_tex) {
            int64_t n = 1;
            for (auto s : tshape) n *= s;
            grad_tex = (n == grad_tex.numel()) ? grad_tex.view(tshape) : at::zeros(tshape, faces.options());      // (empty textures)
        }
        return {grad_faces, grad_tex, at::Tensor(), at::Tensor()};
    }
};

at::Tensor render_faces(const at::Tensor& face_vertices, const at::Tensor& textures, int64_t params_addr, bool anti_aliasing) {
    return RenderFaces::apply(face_vertices, textures, params_addr, anti_aliasing);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "C++ autograd node of gendr_b200.functional.render over the C ABI of libgendr_b200.so";
    m.def("render_faces", &render_faces, "face_vertices [B,F,3,3], textures [B,F,T,3], address of a gendr_render_params struct, anti_aliasing");
    m.def("params_size", []() { return (int64_t)sizeof(gendr_render_params); });
}
