"""A scripted decoder for the temperature-fallback / no-speech / seek logic (tests/test_fallback_ladder.py and
tests/golden/make_fallback_golden.py): what `model.decode` returns for the window that starts at mel frame `seek` at fallback step
`ti`, as a pure function of (scenario, seek, ti), so that the reference's transcribe() and this repo's can be driven by the same
script.  Test infrastructure only."""
import random
import zlib

TB, EOT = 50364, 50257                      # timestamp_begin / eot of the multilingual vocabulary
TEMPERATURES = (0.0, 0.2, 0.4)


def fake_text(tokens):
    """What the scripted tokenizer 'decodes' text tokens to: one character pair per token (repeated tokens -> repetitive text)."""
    return "".join(chr(97 + t % 23) + chr(97 + (t // 23) % 19) for t in tokens if t < EOT)


def compression_ratio(text):                # whisper/utils.py
    data = text.encode("utf-8")
    return len(data) / len(zlib.compress(data)) if data else 0.0


def scripted_result(scenario, seek, ti):
    """-> dict(tokens, avg_logprob, no_speech_prob).  Patterns: timestamp pair + single ending, unfinished tail (the reference seeks
    back), single closing timestamp, no timestamps, empty, an instantaneous segment, repetitive text (compression ratio), low
    log-probability (fallback), silence (no_speech high, logprob low -> no fallback, window skipped)."""
    rng = random.Random(scenario * 1000003 + (seek // 7) * 101 + ti)
    text = lambda n: [rng.randrange(1000, 9000) for _ in range(n)]
    a = rng.randrange(200, 700)              # timestamps in 0.02 s units (window = 1500)
    b = rng.randrange(a + 50, 1400)
    kind = rng.choice(("pair+end", "tail", "tail", "single", "plain", "empty", "instant", "repeat", "pair+end"))
    if kind == "pair+end":
        tokens = [TB] + text(6) + [TB + a, TB + a] + text(5) + [TB + b]
    elif kind == "tail":
        tokens = [TB] + text(6) + [TB + a, TB + a] + text(4)
    elif kind == "single":
        tokens = text(7) + [TB + b]
    elif kind == "plain":
        tokens = text(9)
    elif kind == "empty":
        tokens = []
    elif kind == "instant":
        tokens = [TB] + text(5) + [TB + a, TB + a, TB + a, TB + a] + text(3) + [TB + b]
    else:                                    # repetitive
        tokens = [TB] + [1234] * 80 + [TB + b]
    quality = rng.choice(("good", "good", "low", "silence", "loud-silence"))
    avg_logprob = {"good": -0.35, "low": -1.6, "silence": -1.9, "loud-silence": -0.4}[quality]
    no_speech_prob = {"good": 0.05, "low": 0.2, "silence": 0.93, "loud-silence": 0.9}[quality]
    return {"tokens": tokens, "avg_logprob": avg_logprob, "no_speech_prob": no_speech_prob}
