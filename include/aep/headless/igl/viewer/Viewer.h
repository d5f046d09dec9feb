/* igl/viewer/Viewer.h -- HEADLESS stand-in for libigl's (<= 1.x) viewer, so that the reference's own main.cpp
 * (AnisotropicElastoplasticity/main.cpp, unmodified) builds and runs against the B200 host classes on a machine without
 * OpenGL.  Rendering is out of scope; what main.cpp needs from the viewer is its event loop:
 *   launch() "presses" the key main.cpp starts the simulation on ('s', main.cpp:31-37: a detached thread running
 *   solver.solve(0.3, 100.0, 0.95)), then calls callback_pre_draw at 60 Hz like a render loop would (main.cpp:19-24:
 *   viewer.data.clear(); solver.updateViewer() under the solver's mutex) for AEP_HEADLESS_SECONDS wall seconds
 *   (default 10) and ends the process without running static destructors -- the simulation thread is detached and still
 *   stepping, exactly as when a user closes the reference's window mid-run.
 * The frames are where the reference writes them: particle/particle_N.obj, mesh/mesh_N.obj under the working directory. */
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <thread>
#include "../../../EigenShim.h"

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
    float point_size = 1.0f;
    bool is_animating = false, show_lines = true;
    Eigen::Vector4f background_color;
};
class Viewer {
public:
    ViewerData data;
    ViewerCore core;
    std::function<bool(Viewer&, unsigned char, int)> callback_key_down;
    std::function<bool(Viewer&)> callback_pre_draw;
    int launch() {
        const char* env = std::getenv("AEP_HEADLESS_SECONDS");
        const double seconds = env ? std::atof(env) : 10.0;
        if (callback_key_down) callback_key_down(*this, 's', 0);
        const auto t0 = std::chrono::steady_clock::now();
        long draws = 0;
        while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
            if (callback_pre_draw) callback_pre_draw(*this);
            ++draws;
            std::this_thread::sleep_for(std::chrono::microseconds(16667));
        }
        std::fprintf(stderr, "headless viewer: %ld draw calls in %.1f s, leaving\n", draws, seconds);
        std::fflush(nullptr);
        std::_Exit(0);
    }
};
}  // namespace viewer
}  // namespace igl
