#!/usr/bin/env python
"""calibrate() latency of the reference extractor on one 640x480 image (random weights: there is no checkpoint
offline): reference optimizer + eager CNN / engine + eager CNN / engine + CUDA-graph front end, and the CNN alone."""
import json
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "baseline"))
import ref_import

ref_import.import_reference()
import geocalib
from geocalib.geocalib import GeoCalib as Net

import geocalib_b200 as G
from geocalib_b200 import frontend as FE

torch.manual_seed(7)
with tempfile.TemporaryDirectory() as d:
    torch.save({"model": Net().state_dict()}, Path(d) / "w.tar")
    ex = geocalib.GeoCalib(weights=str(Path(d) / "w.tar")).cuda().eval()
img = torch.rand(3, 480, 640, generator=torch.Generator().manual_seed(3)).cuda()


def wall(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


rec = {"image": [480, 640], "weights": "random", "camera_model": "pinhole"}
rec["calibrate_ms_reference"] = round(wall(lambda: ex.calibrate(img), n=5, warm=2), 3)
G.install(ex)
rec["calibrate_ms_engine_eager_cnn"] = round(wall(lambda: ex.calibrate(img)), 3)
data = ex.image_processor(img[None])
x = data["image"]
rec["cnn_input"] = list(x.shape)
undo = FE.use_device_bases(ex.model)
with torch.no_grad():
    rec["cnn_ms_eager"] = round(wall(lambda: FE.cnn_forward(ex.model, x)), 3)
front_only = FE.GraphedFrontEnd(ex.model)
rec["cnn_ms_graph"] = round(wall(lambda: front_only(x)), 3)
undo()
front = G.install_frontend(ex)
rec["calibrate_ms_engine_graph_cnn"] = round(wall(lambda: ex.calibrate(img)), 3)
rec["graph_captures"], rec["graph_replays"] = front.captures, front.replays
print(json.dumps(rec))
