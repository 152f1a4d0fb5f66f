import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('complexity-tokenizer_b200', 'oracle', 'fixtures', '.'):
    q = os.path.join(ROOT, p)
    if q not in sys.path:
        sys.path.insert(0, q)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def built_lib():
    """libctk.so built in-tree (nvcc cross-compiles without a GPU)."""
    import build as ctk_build
    ctk_build.build()
    import complexity_tokenizer
    return complexity_tokenizer._lib()


@pytest.fixture(scope='session')
def tok_paths():
    import synth
    return {'config1': synth.tokenizer_config1(), 'config2': synth.tokenizer_config2(), 'config3': synth.tokenizer_config3()}


@pytest.fixture(scope='session')
def small_tok_json():
    """A small byte-level tokenizer (600 merges trained on 200 KB), as a JSON string."""
    import json
    import synth
    text, offs = synth.gen_corpus('english', 77, 200 << 10)
    pairs = synth.train_merges(text, 600)
    return json.dumps(synth.assemble_tokenizer(pairs, specials_first=('<unk>', '<pad>', '<s>', '</s>')), ensure_ascii=False)
