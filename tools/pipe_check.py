"""Config-5 pipeline on one replica (bench.pipeline_512) run stand-alone: python tools/pipe_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import mgea_b200 as mg
for _ in range(2):
    r = bench.pipeline_512(mg)
    print("pipeline wall_s %.4f tokens/s %.0f" % (r["wall_s"], r["tokens_per_s"]))
