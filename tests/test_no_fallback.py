"""Guards for the 'no CPU / library fallback' rule: nothing under ai_toolkit_b200/ may import the oracle or route the hot
path through torch compute libraries (SDPA, F.linear, matmul); the oracle is test infrastructure only."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ai_toolkit_b200")


def _sources():
    for fn in sorted(os.listdir(PKG)):
        if fn.endswith(".py"):
            yield fn, open(os.path.join(PKG, fn)).read()


def test_product_never_imports_the_oracle():
    for fn, text in _sources():
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
        assert "oracle." not in re.sub(r'""".*?"""', "", text, flags=re.S), fn


def test_hot_path_has_no_library_compute():
    banned = ("scaled_dot_product_attention", "F.linear(", "torch.matmul(", "torch.bmm(", ".softmax(", "F.layer_norm(", "F.gelu(",
              "torch.compile", "triton")
    for fn, text in _sources():
        code = re.sub(r'""".*?"""', "", text, flags=re.S)
        code = "\n".join(l.split("#")[0] for l in code.splitlines())
        for b in banned:
            assert b not in code, (fn, b)
