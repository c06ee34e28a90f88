import sys, os, time, traceback
sys.path.insert(0, '/root/repo')
import numpy as np
import graphqec_b200 as gq
from graphqec_b200 import backend as be
import itertools
POINTS = [("UnrotatedSurfaceCode", 23, 0.004), ("UnrotatedSurfaceCode", 23, 0.005), ("RotatedSurfaceCode", 23, 0.01), ("UnrotatedSurfaceCode", 20, 0.007)]
for code_name, d, p in POINTS:
    for _ in (0,):
        for _ in (0,):
            t0 = time.time()
            try:
                code = getattr(gq, code_name)(distance=d, depolarize1_rate=p, depolarize2_rate=p)
                code.build_memory_circuit(number_of_rounds=d, logic_check="Z")
                dem = code.memory_circuit.detector_error_model(decompose_errors=True)
                h = be.DemHandle(dem); dec = be.DecoderHandle(h, be.MATCHING)
                gi = dec.graph_info
                res = dec.collect(1, 0, 8192, 0)
                res2 = dec.collect(1, 8192, 1 << 15, 300)
                print(code_name, d, p, "D", dem.num_detectors, "tile", h.tile_shots, "maxdist", gi.max_distance, res, res2, "%.1fs" % (time.time() - t0), flush=True)
                dec.close(); h.close()
            except Exception as e:
                print(code_name, d, p, "FAILED", repr(e)[:200], flush=True)
