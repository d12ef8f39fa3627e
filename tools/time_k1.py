"""Times the game-logic kernel (K1) exactly as bench.py does."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import corintho_ai_b200 as cb
import bench
cb.lib().cb200_set_device(0)
r = bench.measure_game_logic(torch, cb, torch.device("cuda", 0))
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k != "roofline"}, r["roofline"]["frac"])
