"""Golden cases for the word-grouping heuristics.  RUN IN THE BUILD CONTAINER ONLY (imports /root/reference):

    python tests/golden/make_word_golden.py

Feeds synthetic word alignments through the REFERENCE's add_word_timestamps (whisper/timing.py:268-376, find_alignment replaced by
the synthetic alignment) and stores inputs + outputs in tests/golden/ref_words.json for tests/test_host_logic.py."""
import copy
import json
import os
import random
import sys

sys.path.insert(0, "/root/reference")
from whisper import timing as rt                      # noqa: E402  (the reference)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_words.json")
random.seed(0)
POOL = [" hello", " world", ".", ",", " (", ' "', " a", "b", " the", "!", " -", "?", " foo", "bar", '"', ")"]
cases = []
for trial in range(120):
    n, t, tok, al = random.randint(1, 14), 0.0, 0, []
    for _ in range(n):
        ntok = random.randint(1, 3)
        s = t + random.choice([0, 0, 0.3, 2.0]); e = s + random.choice([0.0, 0.1, 0.2, 0.5, 1.5, 3.0]); t = e
        al.append(dict(word=random.choice(POOL), tokens=list(range(tok, tok + ntok)), start=s, end=e, probability=round(random.random(), 4)))
        tok += ntok
    cuts = sorted(random.sample(range(1, tok), min(random.randint(0, 2), tok - 1))) if tok > 1 else []
    bounds = [0] + cuts + [tok]
    segs = [dict(seek=3000, start=round(30 + random.random() * 5, 2), end=round(35 + random.random() * 10, 2),
                 tokens=[50364] + list(range(a, b)) + [50400]) for a, b in zip(bounds[:-1], bounds[1:])]
    last = random.choice([0.0, 28.0, 31.5])
    out = copy.deepcopy(segs)

    class Tk:
        eot = 50257
    rt_find = rt.find_alignment
    rt.find_alignment = lambda *a, **k: [rt.WordTiming(w["word"], list(w["tokens"]), w["start"], w["end"], w["probability"]) for w in al]
    rt.add_word_timestamps(segments=out, model=None, tokenizer=Tk(), num_frames=3000, last_speech_timestamp=last)
    rt.find_alignment = rt_find
    cases.append(dict(alignment=al, segments=segs, last_speech_timestamp=last, expected=out))
with open(OUT, "w") as f:
    json.dump(cases, f)
print("wrote", OUT, len(cases), "cases")
