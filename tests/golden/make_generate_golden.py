"""Generate tests/golden/generate_golden.npz: the token ids and last-step logits of the CPU checker's generation loop (tests/test_generate.py
oracle_generate: embedding -> context decoder -> final RMSNorm -> LM head -> greedy pick -> decode steps, the composition the reference's
src/models/llama/llama.cpp:165-398 intends) on a seeded tiny model.  The fixture pins that loop against drift: tests/test_generate.py replays
it on the CPU (`-m "not gpu"`) and compares the engine's ids with it on the GPU.

    python tests/golden/make_generate_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_generate as tg  # noqa: E402
from test_decoder_engine import make_model  # noqa: E402


def main():
    model = make_model(tg.CFG, seed=tg.GOLDEN["model_seed"], bias=False)
    emb, gamma, lm = tg.tail_weights(tg.GOLDEN["tail_seed"], "f32")
    prompt = np.random.default_rng(tg.GOLDEN["prompt_seed"]).integers(3, tg.V, size=(tg.GOLDEN["batch"], tg.GOLDEN["prompt_len"])).astype(np.int32)
    ids, logits = tg.oracle_generate(model, emb, gamma, lm, prompt, tg.GOLDEN["new_tokens"])
    np.savez_compressed(os.path.join(HERE, "generate_golden.npz"), prompt=prompt, ids=ids, last_logits=logits[-1].astype(np.float32))
    print("ids", ids)


if __name__ == "__main__":
    main()
