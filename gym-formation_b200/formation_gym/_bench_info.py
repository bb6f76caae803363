"""``benchmark_data`` info channel (formation_hd_env.py:97-117, basic_formation_env.py:67-87):
reward, collisions (self included), summed min agent-landmark distance, occupied landmarks.

This is the optional ``make_env(benchmark=True)`` diagnostics hook, not the step path; it reads
the host-side records of one env (O(N*L) scalar work) and takes the reward from the kernel."""
import numpy as np


def benchmark_info(scenario, agent, world, half_threshold):
    rew = scenario.reward(agent, world)
    collisions = 0
    if agent.collide:
        for a in world.agents:
            if scenario.is_collision(a, agent):
                collisions += 1
    P = np.stack([a.state.p_pos for a in world.agents])
    min_dists, occupied = 0.0, 0
    for l in world.landmarks:
        d = np.sqrt(np.sum(np.square(P - l.state.p_pos), axis=1)).min()
        min_dists += d
        occupied += int(d < 0.1)
    return {'reward': rew, 'collisions': collisions, 'min_dists': min_dists,
            'occupied_landmarks': occupied}
