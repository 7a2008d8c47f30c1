import sys; sys.path.insert(0,"multimodal-rssm_b200")
import os, torch
from mrssm_b200 import data
c = data._device_local_cpus("cuda:0")
print("local cpus:", None if c is None else (min(c), max(c), len(c)), "affinity:", len(os.sched_getaffinity(0)))
