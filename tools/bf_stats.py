# debug helper: one A4 page through the engine with the blackfilter built with -DBF_STATS
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import unpaper_gpu_b200 as U
from unpaper_gpu_b200 import synth
from unpaper_gpu_b200.lib import Engine
import torch
w, h = synth.A4_W, synth.A4_H
pages = synth.gray_page(0, w, h)[None]
eng = Engine(U.default_sheet_config(), w, h, U.FMT_GRAY8, group_pages=1, lanes=1)
out, res = eng.process_numpy(pages)
torch.cuda.synchronize()
print("fills", res[0].blackfilter_fills)
