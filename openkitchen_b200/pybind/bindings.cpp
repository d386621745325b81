// bindings.cpp -- python module `open_kitchen_pybind`: same class / method names as the reference's
// Pybind/bindings.cpp:62-79 (Environment(race_track_path, draw_rays, hidden_window), set_action, step,
// get_render_target, get_render_target_info; RenderTargetInfo{width,height,channels,row_bytes()}), served by
// the B200 shim.  The reference exposes only the 8.96 MB render target; here the state a lidar policy needs
// is reachable too (pose, lidar, flags, reset).  Batch / zero-copy users: openkitchen_b200.BatchEnv.
#include "Environment/Environment.h"

#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <memory>
#include <tuple>
#include <vector>

namespace py = pybind11;

namespace
{
class ScriptedAgent : public Agent
{
  public:
    using Agent::Agent;
    void updateAction() override {} // the action is written from Python
};

class PyEnv
{
  public:
    PyEnv(const std::string &race_track_path, bool draw_rays, bool hidden_window)
    {
        agents_.push_back(std::make_unique<ScriptedAgent>(Vec2d{0.F, 0.F}, 0.F, 0));
        env_ = std::make_unique<Environment>(race_track_path, createBaseAgentPtrs(agents_), draw_rays, hidden_window);
        env_->resetAgent(agents_[0].get(), /*pick_random_point=*/true); // bindings.cpp:27-32 starts at a random index
        env_->visualizer_->setAgentToFollow(agents_[0].get());
    }
    void set_action(float throttle_delta, float steering_delta)
    {
        agents_[0]->current_action_.throttle_delta = throttle_delta;
        agents_[0]->current_action_.steering_delta = steering_delta;
    }
    void step() { env_->step(); }
    std::vector<uint8_t>            render_target() { return env_->getRenderTargetHost(); }
    ScreenGrabber::RenderTargetInfo render_target_info() { return env_->getRenderTargetInfo(); }

    void seed(uint64_t s) { env_->seed(s); }
    void reset(bool random_point, bool lane, bool heading) { env_->resetAgent(agents_[0].get(), random_point, lane, heading); }
    void set_movement_mode(int mode) { agents_[0]->setMovementMode(static_cast<Agent::MovementMode>(mode)); }
    std::tuple<float, float, float, float> pose() const
    {
        const Agent &a = *agents_[0];
        return {a.pos_.x, a.pos_.y, a.rot_, a.speed_};
    }
    std::vector<std::pair<float, float>> lidar() const
    {
        std::vector<std::pair<float, float>> out;
        for (const Vec2d &h : agents_[0]->sensor_hits_)
            out.emplace_back(h.x, h.y);
        return out;
    }
    bool   crashed() const { return agents_[0]->crashed_; }
    bool   timed_out() const { return agents_[0]->timed_out_; }
    size_t nearest_track_index() const { return env_->race_track_->findNearestTrackIndexBruteForce(agents_[0]->pos_); }

  private:
    std::vector<std::unique_ptr<ScriptedAgent>> agents_;
    std::unique_ptr<Environment>                env_;
};
} // namespace

PYBIND11_MODULE(open_kitchen_pybind, m)
{
    m.doc() = "OpenKitchen racing environment on the B200 step kernels (drop-in for the reference's Pybind module)";
    py::class_<ScreenGrabber::RenderTargetInfo>(m, "RenderTargetInfo")
        .def_readonly("width", &ScreenGrabber::RenderTargetInfo::width)
        .def_readonly("height", &ScreenGrabber::RenderTargetInfo::height)
        .def_readonly("channels", &ScreenGrabber::RenderTargetInfo::channels)
        .def("row_bytes", &ScreenGrabber::RenderTargetInfo::row_bytes);
    py::class_<PyEnv>(m, "Environment")
        .def(py::init<const std::string &, bool, bool>(), py::arg("race_track_path"), py::arg("draw_rays") = true,
             py::arg("hidden_window") = true)
        .def("set_action", &PyEnv::set_action)
        .def("step", &PyEnv::step, py::call_guard<py::gil_scoped_release>())
        .def("get_render_target", &PyEnv::render_target)
        .def("get_render_target_info", &PyEnv::render_target_info)
        // beyond the reference
        .def("seed", &PyEnv::seed)
        .def("reset", &PyEnv::reset, py::arg("random_point") = true, py::arg("randomize_lane") = false,
             py::arg("randomize_heading") = false)
        .def("set_movement_mode", &PyEnv::set_movement_mode)
        .def("pose", &PyEnv::pose)
        .def("lidar", &PyEnv::lidar)
        .def_property_readonly("crashed", &PyEnv::crashed)
        .def_property_readonly("timed_out", &PyEnv::timed_out)
        .def("nearest_track_index", &PyEnv::nearest_track_index);
}
