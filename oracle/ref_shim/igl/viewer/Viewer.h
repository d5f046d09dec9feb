// Stand-in for libigl's (<= 1.x) igl/viewer/Viewer.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// The reference's simulation sources include the viewer header for three things: igl::PI, the stream headers it drags in,
// and the Viewer type whose `data` member they push render buffers into.  The headless build of the reference
// (oracle/Makefile, _ref/libaep_ref.so) never draws, so every viewer call is a no-op here.
#pragma once
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <Eigen/Core>

namespace igl {
const double PI = 3.1415926535897932384626433832795;
namespace viewer {
struct ViewerData {
    void clear() {}
    template <class... A> void add_points(const A&...) {}
    template <class... A> void set_points(const A&...) {}
    template <class... A> void add_edges(const A&...) {}
    template <class... A> void set_mesh(const A&...) {}
    template <class... A> void set_colors(const A&...) {}
};
struct ViewerCore {
    double point_size = 1.0; bool is_animating = false, show_lines = true;
    Eigen::Vector4f background_color;
};
class Viewer {
public:
    ViewerData data;
    ViewerCore core;
    std::function<bool(Viewer&, unsigned char, int)> callback_key_down;
    std::function<bool(Viewer&)> callback_pre_draw;
    int launch() { return 0; }
};
}  // namespace viewer
}  // namespace igl
