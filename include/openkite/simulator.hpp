// simulator.hpp -- ROS-free host mirror of openKITE's Simulator stepping loop and state / pose record formats
// (reference: src/kite_model/simulator.{h,cpp}: simulate() :43-51, state field order :59-74, pose = state[6:13] :21;
// text log lines of src/nodes/simple_logger.cpp:63-85).  The ROS transport itself (topics, callbacks, rates) is out of
// scope; what callers of the hot path need from it -- "advance the state by tf with the held control" and "write the
// 13-state vector in the simulator / logger field order" -- is kept, on top of the GPU-backed ODESolver.
#pragma once
#include <iomanip>
#include <ostream>

#include "integrator.hpp"

namespace openkite {

class Simulator {
public:
    explicit Simulator(const ODESolver& object) : m_object(std::make_shared<ODESolver>(object)), controls(DM::zeros(3)), initialized(false) {}
    virtual ~Simulator() {}

    /** one step of length params["tf"] with the current controls (simulator.cpp:43-51) */
    void simulate() {
        Dict p = m_object->getParams();
        const double dt = p["tf"];
        state = m_object->solve(state, controls, dt);
    }
    DM getState() { return state; }
    DM getPose() { DM p(7, 1); for (int i = 0; i < 7; ++i) p[i] = state[6 + i]; return p; }     // simulator.h:21
    bool is_initialized() { return initialized; }
    void initialize(const DM& init_value) { state = init_value; initialized = true; }
    /** what controlCallback stores (simulator.cpp:27-33): thrust, elevator, rudder */
    void setControls(double thrust, double elevator, double rudder) { controls = DM{thrust, elevator, rudder}; }

    /** "/kite_state" record in the logger's text format (simple_logger.cpp:73-85): stamp, twist linear (v), twist angular
     *  (w), translation (r), rotation w x y z (q) -- i.e. the state vector in its native order, fixed, 8 decimals. */
    void write_state(std::ostream& os, double stamp) const { write_record(os, stamp, state.ptr(), 13); }
    /** pose record (simple_logger.cpp:63-68): stamp, position, orientation w x y z */
    void write_pose(std::ostream& os, double stamp) const { write_record(os, stamp, state.ptr() + 6, 7); }
    static void write_record(std::ostream& os, double stamp, const double* v, int n) {
        os << std::fixed << std::setprecision(8) << stamp << " ";
        for (int i = 0; i < n; ++i) os << v[i] << (i + 1 < n ? " " : "\n");
    }

private:
    std::shared_ptr<ODESolver> m_object;
    DM controls, state;
    bool initialized;
};

}  // namespace openkite
