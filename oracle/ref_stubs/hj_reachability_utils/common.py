"""Stand-in for `hj_reachability_utils.common` (un-pinned, un-vendored git repo
ChoiJangho/hj_reachability_utils; reference README.md:38-41). TEST INFRASTRUCTURE ONLY.

Used by the reference at multiagent/safety_filter.py:6-7,15,85,163 and
multiagent/custom_scenarios/navigation_graph_safe.py:24,133-138.
"""
import numpy as np
import hj_reachability as hj


class GridMetaData(object):
    def __init__(self, domain_lo, domain_hi, shape, periodic_dims=()):
        self.domain_lo = np.asarray(domain_lo, dtype=np.float64)
        self.domain_hi = np.asarray(domain_hi, dtype=np.float64)
        self.shape = tuple(int(s) for s in shape)
        self.periodic_dims = tuple(int(d) for d in periodic_dims)


def get_hj_grid_from_meta_data(meta):
    return hj.Grid(meta.domain_lo, meta.domain_hi, meta.shape, meta.periodic_dims)


class HjData(object):
    """Pickled value-function container: `.info['separation_distance']`, `.grid_meta_data`, `.values`."""

    def __init__(self, values, grid_meta_data, separation_distance):
        self.values = values
        self.grid_meta_data = grid_meta_data
        self.info = {'separation_distance': float(separation_distance)}


class TtrData(object):
    """Pickled time-to-reach container: `.grid_meta_data`, `.values`, `.ttr_max`."""

    def __init__(self, values, grid_meta_data, ttr_max):
        self.values = values
        self.grid_meta_data = grid_meta_data
        self.ttr_max = float(ttr_max)


class ControlAndDisturbanceAffineDynamics(object):
    """x' = f(x,t) + G_u(x,t) u + G_d(x,t) d  (hj_reachability.dynamics of the same name)."""

    def __init__(self, control_mode, disturbance_mode, control_space, disturbance_space):
        self.control_mode = control_mode
        self.disturbance_mode = disturbance_mode
        self.control_space = control_space
        self.disturbance_space = disturbance_space

    def __call__(self, state, control, disturbance, time):
        return (self.open_loop_dynamics(state, time)
                + self.control_jacobian(state, time) @ control
                + self.disturbance_jacobian(state, time) @ disturbance)

    def optimal_control_and_disturbance(self, state, time, grad_value):
        control_direction = grad_value @ self.control_jacobian(state, time)
        if self.control_mode == "min":
            control_direction = -control_direction
        disturbance_direction = grad_value @ self.disturbance_jacobian(state, time)
        if self.disturbance_mode == "min":
            disturbance_direction = -disturbance_direction
        return (self.control_space.extreme_point(control_direction),
                self.disturbance_space.extreme_point(disturbance_direction))

    def optimal_control(self, state, time, grad_value):
        return self.optimal_control_and_disturbance(state, time, grad_value)[0]

    def optimal_disturbance(self, state, time, grad_value):
        return self.optimal_control_and_disturbance(state, time, grad_value)[1]
