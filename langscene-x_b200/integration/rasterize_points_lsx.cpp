// rasterize_points_lsx.cpp — INTEGRATION.md "Route B" as a compilable file: the reference's own C++ glue
// (diff-langsurf-rasterizer/rasterize_points.{h,cu}, ext.cpp; simple-knn/spatial.{h,cu}, ext.cpp) re-bound on the C ABI of
// liblsx_b200.so.  Same pybind names, same argument lists, same return tuples as the reference's `_C` modules, so
// diff_LangSurf_rasterization/__init__.py and `from simple_knn._C import distCUDA2` work unchanged on top of it.
//
// torch types stay on THIS side of the boundary: the library sees raw device pointers, sizes, a stream and three
// allocation callbacks (the C form of resizeFunctional, rasterize_points.cu:27-33).
//
// Build (what tests/test_route_b.py and __graft_entry__.build() do):
//   g++ -O2 -std=c++17 -fPIC -shared rasterize_points_lsx.cpp -I<repo>/include $(torch include flags)
//       -DTORCH_EXTENSION_NAME=lsx_route_b -L<dir of liblsx_b200.so> -llsx_b200 -Wl,-rpath,'$ORIGIN' -o lsx_route_b.so
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <tuple>

#include "lsx_rasterizer.h"

namespace {

char* resize_cb(void* user, size_t bytes) {  // == resizeFunctional, as a C callback
    auto* t = static_cast<torch::Tensor*>(user);
    t->resize_({(long long)bytes});
    return reinterpret_cast<char*>(t->contiguous().data_ptr());
}

// empty tensor == "absent" (forward.cu:205,241); the caller keeps `keep` alive across the library call
const float* fptr(const torch::Tensor& t, std::vector<torch::Tensor>& keep) {
    if (t.numel() == 0) return nullptr;
    keep.push_back(t.contiguous());
    return keep.back().data_ptr<float>();
}

void* current_stream() { return c10::cuda::getCurrentCUDAStream().stream(); }

}  // namespace

// rasterize_points.cu:35-143
std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                       const torch::Tensor& language_feature, const torch::Tensor& language_feature_instance,
                       const torch::Tensor& opacity, const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& all_map,
                       const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
                       const float tan_fovy, const int image_height, const int image_width, const torch::Tensor& sh,
                       const int degree, const torch::Tensor& campos, const bool prefiltered, const bool render_geo,
                       const bool debug, const bool include_feature) {
    if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
    const int P = means3D.size(0), H = image_height, W = image_width;
    const int F = include_feature ? (int)language_feature.size(1) : 0;            // run-time width (config.h:16 upstream)
    const int Fi = include_feature ? (int)language_feature_instance.size(1) : 0;
    auto float_opts = means3D.options().dtype(torch::kFloat32);
    auto int_opts = means3D.options().dtype(torch::kInt32);
    // torch::empty instead of torch::full(0): the library writes every element
    torch::Tensor out_color = torch::empty({3, H, W}, float_opts);
    torch::Tensor out_language_feature = include_feature ? torch::empty({F, H, W}, float_opts) : torch::zeros({1}, float_opts);
    torch::Tensor out_language_feature_instance =
        include_feature ? torch::empty({Fi, H, W}, float_opts) : torch::zeros({1}, float_opts);
    torch::Tensor radii = torch::empty({P}, int_opts), out_observe = torch::empty({P}, int_opts);
    torch::Tensor out_all_map = torch::empty({5, H, W}, float_opts), out_plane_depth = torch::empty({1, H, W}, float_opts);
    auto byte_opts = means3D.options().dtype(torch::kByte);
    torch::Tensor geomBuffer = torch::empty({0}, byte_opts), binningBuffer = torch::empty({0}, byte_opts),
                  imgBuffer = torch::empty({0}, byte_opts);

    std::vector<torch::Tensor> keep;
    keep.reserve(16);
    lsx_forward_args a{};
    a.P = P; a.D = degree; a.M = (sh.numel() != 0 && sh.dim() >= 2) ? (int)sh.size(1) : 0; a.W = W; a.H = H; a.F = F; a.Fi = Fi;
    a.tanfovx = tan_fovx; a.tanfovy = tan_fovy; a.scale_modifier = scale_modifier;
    a.prefiltered = prefiltered; a.render_geo = render_geo; a.debug = debug; a.include_feature = include_feature;
    a.background = fptr(background, keep); a.means3D = fptr(means3D, keep); a.shs = fptr(sh, keep);
    a.colors_precomp = fptr(colors, keep);
    a.language_feature = include_feature ? fptr(language_feature, keep) : nullptr;
    a.language_feature_instance = include_feature ? fptr(language_feature_instance, keep) : nullptr;
    a.opacities = fptr(opacity, keep); a.scales = fptr(scales, keep); a.rotations = fptr(rotations, keep);
    a.cov3D_precomp = fptr(cov3D_precomp, keep); a.all_map = fptr(all_map, keep);
    a.viewmatrix = fptr(viewmatrix, keep); a.projmatrix = fptr(projmatrix, keep); a.campos = fptr(campos, keep);
    a.out_color = out_color.data_ptr<float>();
    a.out_language_feature = include_feature ? out_language_feature.data_ptr<float>() : nullptr;
    a.out_language_feature_instance = include_feature ? out_language_feature_instance.data_ptr<float>() : nullptr;
    a.radii = P ? radii.data_ptr<int>() : nullptr; a.out_observe = P ? out_observe.data_ptr<int>() : nullptr;
    a.out_all_map = out_all_map.data_ptr<float>(); a.out_plane_depth = out_plane_depth.data_ptr<float>();
    a.geom_alloc = resize_cb; a.geom_user = &geomBuffer;
    a.binning_alloc = resize_cb; a.binning_user = &binningBuffer;
    a.image_alloc = resize_cb; a.image_user = &imgBuffer;
    a.stream = current_stream();
    int32_t rendered = 0;
    if (lsx_rasterize_forward(&a, &rendered) != 0) AT_ERROR(lsx_last_error());
    return std::make_tuple((int)rendered, out_color, out_language_feature, out_language_feature_instance, radii, out_observe,
                           out_all_map, out_plane_depth, geomBuffer, binningBuffer, imgBuffer);
}

// rasterize_points.cu:145-259
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& all_map_pixels,
                               const torch::Tensor& means3D, const torch::Tensor& radii, const torch::Tensor& colors,
                               const torch::Tensor& language_feature, const torch::Tensor& language_feature_instance,
                               const torch::Tensor& all_maps, const torch::Tensor& scales, const torch::Tensor& rotations,
                               const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                               const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                               const torch::Tensor& dL_dout_color, const torch::Tensor& dL_dout_language_feature,
                               const torch::Tensor& dL_dout_language_feature_instance, const torch::Tensor& dL_dout_all_map,
                               const torch::Tensor& dL_dout_plane_depth, const torch::Tensor& sh, const int degree,
                               const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer, const bool render_geo,
                               const bool debug, const bool include_feature) {
    const int P = means3D.size(0), H = dL_dout_color.size(1), W = dL_dout_color.size(2);
    const int M = (sh.numel() != 0 && sh.dim() >= 2) ? (int)sh.size(1) : 0;
    const int F = include_feature ? (int)language_feature.size(1) : 0;
    const int Fi = include_feature ? (int)language_feature_instance.size(1) : 0;
    auto o = means3D.options().dtype(torch::kFloat32);
    // torch::empty instead of torch::zeros: the library writes every row (zeros for culled splats)
    torch::Tensor dL_dmeans3D = torch::empty({P, 3}, o), dL_dmeans2D = torch::empty({P, 3}, o),
                  dL_dmeans2D_abs = torch::empty({P, 3}, o), dL_dcolors = torch::empty({P, 3}, o);
    torch::Tensor dL_dlanguage_feature = include_feature ? torch::empty({P, F}, o) : torch::zeros({1}, o);
    torch::Tensor dL_dlanguage_feature_instance = include_feature ? torch::empty({P, Fi}, o) : torch::zeros({1}, o);
    torch::Tensor dL_dall_map = torch::empty({P, 5}, o), dL_dconic = torch::empty({P, 2, 2}, o),
                  dL_dopacity = torch::empty({P, 1}, o), dL_dcov3D = torch::empty({P, 6}, o),
                  dL_dsh = torch::empty({P, M, 3}, o), dL_dscales = torch::empty({P, 3}, o),
                  dL_drotations = torch::empty({P, 4}, o);
    if (P != 0) {
        std::vector<torch::Tensor> keep;
        keep.reserve(24);
        lsx_backward_args a{};
        a.P = P; a.D = degree; a.M = M; a.W = W; a.H = H; a.F = F; a.Fi = Fi; a.R = R;
        a.tanfovx = tan_fovx; a.tanfovy = tan_fovy; a.scale_modifier = scale_modifier;
        a.render_geo = render_geo; a.debug = debug; a.include_feature = include_feature;
        a.background = fptr(background, keep); a.means3D = fptr(means3D, keep); a.shs = fptr(sh, keep);
        a.colors_precomp = fptr(colors, keep);
        a.language_feature = include_feature ? fptr(language_feature, keep) : nullptr;
        a.language_feature_instance = include_feature ? fptr(language_feature_instance, keep) : nullptr;
        a.all_map = fptr(all_maps, keep); a.scales = fptr(scales, keep); a.rotations = fptr(rotations, keep);
        a.cov3D_precomp = fptr(cov3D_precomp, keep);
        a.viewmatrix = fptr(viewmatrix, keep); a.projmatrix = fptr(projmatrix, keep); a.campos = fptr(campos, keep);
        keep.push_back(radii.contiguous());
        a.radii = keep.back().data_ptr<int>();
        a.out_all_map = fptr(all_map_pixels, keep);
        a.geom_buffer = reinterpret_cast<const char*>(geomBuffer.contiguous().data_ptr());
        a.binning_buffer = binningBuffer.numel() ? reinterpret_cast<const char*>(binningBuffer.contiguous().data_ptr()) : nullptr;
        a.image_buffer = reinterpret_cast<const char*>(imageBuffer.contiguous().data_ptr());
        a.dL_dout_color = fptr(dL_dout_color, keep);
        a.dL_dout_language_feature = include_feature ? fptr(dL_dout_language_feature, keep) : nullptr;
        a.dL_dout_language_feature_instance = include_feature ? fptr(dL_dout_language_feature_instance, keep) : nullptr;
        a.dL_dout_all_map = fptr(dL_dout_all_map, keep); a.dL_dout_plane_depth = fptr(dL_dout_plane_depth, keep);
        a.dL_dmeans2D = dL_dmeans2D.data_ptr<float>(); a.dL_dmeans2D_abs = dL_dmeans2D_abs.data_ptr<float>();
        a.dL_dconic = dL_dconic.data_ptr<float>(); a.dL_dopacity = dL_dopacity.data_ptr<float>();
        a.dL_dcolors = dL_dcolors.data_ptr<float>();
        a.dL_dlanguage_feature = include_feature ? dL_dlanguage_feature.data_ptr<float>() : nullptr;
        a.dL_dlanguage_feature_instance = include_feature ? dL_dlanguage_feature_instance.data_ptr<float>() : nullptr;
        a.dL_dmeans3D = dL_dmeans3D.data_ptr<float>(); a.dL_dcov3D = dL_dcov3D.data_ptr<float>();
        a.dL_dsh = M > 0 ? dL_dsh.data_ptr<float>() : nullptr;
        a.dL_dscales = dL_dscales.data_ptr<float>(); a.dL_drotations = dL_drotations.data_ptr<float>();
        a.dL_dall_map = dL_dall_map.data_ptr<float>();
        a.stream = current_stream();
        a.accumulate_param_grads = 0;
        a.binning_bytes = (uint64_t)binningBuffer.numel();   // the list capacity is recovered from the buffer's size
        if (lsx_rasterize_backward(&a) != 0) AT_ERROR(lsx_last_error());
    }
    return std::make_tuple(dL_dmeans2D, dL_dmeans2D_abs, dL_dcolors, dL_dlanguage_feature, dL_dlanguage_feature_instance,
                           dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, dL_dall_map);
}

// rasterize_points.cu:261-280
torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix) {
    const int P = means3D.size(0);
    torch::Tensor present = torch::empty({P}, means3D.options().dtype(at::kBool));
    if (P != 0) {
        auto m = means3D.contiguous(), v = viewmatrix.contiguous(), pr = projmatrix.contiguous();
        if (lsx_mark_visible(P, m.data_ptr<float>(), v.data_ptr<float>(), pr.data_ptr<float>(),
                             reinterpret_cast<uint8_t*>(present.data_ptr<bool>()), current_stream()) != 0)
            AT_ERROR(lsx_last_error());
    }
    return present;
}

// simple-knn/spatial.cu:15-25
torch::Tensor distCUDA2(const torch::Tensor& points) {
    const int P = points.size(0);
    torch::Tensor means = torch::empty({P}, points.options().dtype(torch::kFloat32));
    if (P != 0) {
        torch::Tensor scratch = torch::empty({0}, points.options().dtype(torch::kByte));
        auto pts = points.contiguous();
        if (lsx_knn_mean_dist2(P, pts.data_ptr<float>(), means.data_ptr<float>(), resize_cb, &scratch, current_stream()) != 0)
            AT_ERROR(lsx_last_error());
    }
    return means;
}

// ext.cpp of both submodules, in one module
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
    m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
    m.def("mark_visible", &markVisible);
    m.def("distCUDA2", &distCUDA2);
}
