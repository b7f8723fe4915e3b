#!/usr/bin/env python
"""Golden fixture for SURVEY 8f row N4: the deterministic initial states of the reference's evaluation scenarios,
recorded by building the UNMODIFIED reference scenario class
(/root/reference/multiagent/custom_scenarios/navigation_graph_safe_eval.py: scenario_circular_config :100-121,
scenario_three_vehicle_conflicting_example :320-381, scenario_two_vehicle_conflicting_example :383-431) through
oracle/ref_harness.py's import stubs. -> tests/golden/aux/eval_scenarios.npz   (build container only)"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

CASES = (('circular', 'circular_config', 6, 'double_integrator', 4.0), ('circular_at', 'circular_config', 5, 'airtaxi', 6.0),
         ('two_vehicle_conflict', 'two_vehicle_conflicting_example', 2, 'airtaxi', 6.0),
         ('three_vehicle_conflict', 'three_vehicle_conflicting_example', 3, 'airtaxi', 6.0))


def main():
    H.setup_reference()
    import multiagent.config as C
    out = {}
    for tag, typ, N, dyn, ws in CASES:
        C.eval_scenario_type = typ                      # the scenario module reads it at import time
        import multiagent.custom_scenarios.navigation_graph_safe_eval as EV
        importlib.reload(EV)
        sc = EV.Scenario()
        args = H.make_args(num_agents=N, dynamics_type=dyn, world_size=ws, num_landmarks=sc.get_default_landmark_num_for_scenario())
        np.random.seed(0)
        world = sc.make_world(args)
        out[f'{tag}__agent_values'] = np.array([a.state.values for a in world.agents], dtype=np.float64)
        out[f'{tag}__goal0_pos'] = np.array([world.landmarks[i].state.p_pos for i in range(N)], dtype=np.float64)
        out[f'{tag}__goal0_heading'] = np.array([world.landmarks[i].heading for i in range(N)], dtype=np.float64)
        out[f'{tag}__goal0_speed'] = np.array([world.landmarks[i].speed for i in range(N)], dtype=np.float64)
        out[f'{tag}__world_size'] = np.float64(ws)
    path = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'aux', 'eval_scenarios.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    main()
