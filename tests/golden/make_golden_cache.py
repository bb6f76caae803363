"""Golden fixture for World.cache_dists (core.py:132,156-180,224-225,298-301), build container only, unmodified
reference: formation_hd_env with 5 agents, cache primed by calculate_distances() after the reset, three ordinary steps,
then agent 1 is TELEPORTED onto agent 0's neighbourhood between two steps -- the next step's contact forces still come
from the stale cache (no contact), the one after sees the contact.      python tests/golden/make_golden_cache.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_harness as rh  # noqa: E402


def run(seed=9, n=5, cache=True):
    np.random.seed(seed)
    env = rh.make_reference_env("formation_hd_env", n, 25)
    env.reset()
    w = env.world
    w.cache_dists = cache
    w.calculate_distances()
    sc = rh.scenario_of(env)
    rng = np.random.default_rng(seed)
    d = dict(pos0=np.stack([a.state.p_pos for a in w.agents]), vel0=np.stack([a.state.p_vel for a in w.agents]),
             lm0=np.stack([l.state.p_pos for l in w.landmarks]), shape=np.array(sc.ideal_shape), ivel=np.array(sc.ideal_vel),
             vect0=w.cached_dist_vect.copy(), mag0=w.cached_dist_mag.copy(), mind=w.min_dists.copy(),
             coll0=w.cached_collisions.copy())
    acts, pos, vel, rew, obs = [], [], [], [], []
    for t in range(6):
        if t == 3:                                      # teleport agent 1 next to agent 0 (inside the contact range)
            w.agents[1].state.p_pos = w.agents[0].state.p_pos + np.array([0.04, 0.01])
            d["tele_pos"] = np.stack([a.state.p_pos for a in w.agents])
        a = rng.uniform(-1, 1, (n, 2))
        obs_n, reward_n, done_n, info_n = env.step([x.copy() for x in a])
        acts.append(a); pos.append(np.stack([ag.state.p_pos for ag in w.agents]))
        vel.append(np.stack([ag.state.p_vel for ag in w.agents])); rew.append(reward_n[0][0]); obs.append(np.stack(obs_n))
    d.update(acts=np.stack(acts), pos=np.stack(pos), vel=np.stack(vel), reward=np.array(rew), obs=np.stack(obs),
             vect_last=w.cached_dist_vect.copy(), mag_last=w.cached_dist_mag.copy(), coll_last=w.cached_collisions.copy())
    return d


if __name__ == "__main__":
    d = run()
    fresh = run(cache=False)
    # the fixture exercises the quirk: identical until the teleport, then the stale cache misses the new contact
    assert np.array_equal(d["vel"][:3], fresh["vel"][:3])
    gap = np.abs(d["vel"][3] - fresh["vel"][3]).max()
    print("velocity difference stale vs fresh distances at the teleport step: %.3f" % gap)
    assert gap > 0.05
    np.savez_compressed(os.path.join(HERE, "cache_dists_n5.npz"), **d)
    print({k: v.shape for k, v in d.items()})
