"""Shared test helpers: golden loading and oracle drivers (oracle/ is test infrastructure)."""
import math
import os

import numpy as np

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SEARCH_CASES = ["ttt", "lunar", "breakout", "atari18", "atari18_nonoise_bounds", "flat_ties"]
REPLAY_CASES = ["breakout", "lunar_td1000", "ttt", "breakout_clip"]


def is_clip_case(g):
  """The `breakout_clip` fixture stores RAW rewards; the reference saw them through ClipRewardEnv.reward."""
  return int(g.get("clip_rewards", 0)) == 1


def load(name):
  return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def search_case_cfg(g):
  kb = [None if math.isnan(b) else float(b) for b in g["known_bounds"]]
  return dict(num_simulations=int(g["num_simulations"]), action_space=int(g["action_space"]),
              two_players=bool(g["two_players"]), discount=float(g["discount"]),
              pb_c_base=float(g["pb_c_base"]), pb_c_init=float(g["pb_c_init"]),
              init_value_score=float(g["init_value_score"]), known_bounds=kb)


def oracle_search_from_golden(g, **kw):
  cfg = oracle.make_cfg(sum_mode=int(g["py_sum_mode"]), **search_case_cfg(g))
  hv = g["hashnet"]
  hn = oracle.HashNet(float(hv[0]), float(hv[1]), float(hv[2]), int(hv[3]))
  noise = g["noise"] if int(g["use_noise"]) else None
  return oracle.search(cfg, g["root_logits"], legal_mask=g["legal"], noise=noise,
                       noise_frac=float(g["noise_frac"]), root_to_play=g["to_play"], hashnet=hn,
                       root_state=g["root_state"].astype(np.uint64), dump_tree=True, **kw)


SEARCH_EXACT_KEYS = ["visits", "trace_parent", "trace_action", "trace_depth", "edge_visit",
                     "edge_child", "root_value", "root_vsum", "minmax", "edge_prior", "edge_vsum",
                     "edge_reward"]
