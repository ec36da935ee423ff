"""One rank of tests/test_fallback_gpu.py::test_speculative_seek_two_ranks (launched by torch.distributed.run)."""
import os, sys, json, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth
from tests._util import exported
from whisper_b200.model import ModelDimensions, WhisperB200
from whisper_b200.transcribe import transcribe
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
dims, ckpt, folder = exported("nano", 1, 0.03)
m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
audio = torch.cat([synth.noise_audio(10 + i, 480000) for i in range(4)])[:1600000]
res = transcribe(m, audio, beam_size=5, sample_len=40, seek_mode="reference", window_batch=2, rank=rank, world_size=world)
m.close()
print("RESULT%d" % rank + json.dumps({"seeks": res["seeks"], "tokens": [s["tokens"] for s in res["segments"]]}), flush=True)
dist.destroy_process_group()
