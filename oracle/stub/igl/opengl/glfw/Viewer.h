#pragma once
// TEST INFRASTRUCTURE (oracle build only): headless stand-in for igl::opengl::glfw::Viewer
// (no GL / GLFW in this image).  Only the members main.cpp / Diagnostic.hpp touch exist.
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <igl/colormap.h>
#include <igl/writeOBJ.h>
#include <igl/readOBJ.h>
#include <igl/per_vertex_normals.h>
#include <functional>
namespace igl { namespace opengl {
struct ViewerData {
    Eigen::MatrixXd V, V_uv; Eigen::MatrixXi F;
    bool show_lines = true, show_texture = false, show_overlay = true; float point_size = 1.f;
    void set_edges(const Eigen::MatrixXd&, const Eigen::MatrixXi&, const Eigen::MatrixXd&) {}
    void add_edges(const Eigen::MatrixXd&, const Eigen::MatrixXd&, const Eigen::MatrixXd&) {}
    void set_points(const Eigen::MatrixXd&, const Eigen::MatrixXd&) {}
    void add_points(const Eigen::MatrixXd&, const Eigen::MatrixXd&) {}
    void set_colors(const Eigen::MatrixXd&) {}
    void set_mesh(const Eigen::MatrixXd& v, const Eigen::MatrixXi& f) { V = v; F = f; }
    void set_uv(const Eigen::MatrixXd& uv) { V_uv = uv; }
    void compute_normals() {}
    void clear() { V.resize(0,3); F.resize(0,3); V_uv.resize(0,2); }
};
struct ViewerCore {
    Eigen::Vector4f background_color, viewport = Eigen::Vector4f(0,0,1280,800);
    float model_zoom = 1.f, lighting_factor = 0.f, camera_zoom = 1.f; double animation_max_fps = 60.0;
    bool is_animating = false, orthographic = false; Eigen::Quaternionf trackball_angle;
    void align_camera_center(const Eigen::MatrixXd&, const Eigen::MatrixXi&) {}
    template <class M> void draw_buffer(ViewerData&, bool, M&, M&, M&, M&) {}
};
namespace glfw {
struct Viewer {
    ViewerCore core; ViewerData data_;
    ViewerData& data() { return data_; }
    std::function<bool(Viewer&, unsigned char, int)> callback_key_down;
    std::function<bool(Viewer&)> callback_pre_draw, callback_post_draw;
    int launch() { return 0; }
    int launch_init(bool = true, bool = false) { return 0; }
    bool launch_rendering(bool = true) { return true; }
};
}}}
