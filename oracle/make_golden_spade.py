"""Pins oracle/spade_oracle.py against the UNMODIFIED reference Stage-II generators and writes the golden fixtures.
Run in the build container (needs /root/reference and torchvision):  python oracle/make_golden_spade.py

  tests/golden/spade_keys.json   state_dict keys and shapes of Generator / Generator_audio (drop-in compatibility pin)
  tests/golden/spade_gen.npz     inputs + reference output of Generator        at 64x64 and 96x128, layer outputs at 64x64
  tests/golden/spade_audio.npz   inputs + reference output of Generator_audio  at 64x64
The weights are regenerated from the seed (tests/spade_fixtures.py); the reference modules load them with strict=True."""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "tests")]
REF = os.environ.get("SAHS_REFERENCE", "/root/reference/nerf-pytorch")


def ref_module():
    spec = importlib.util.spec_from_file_location("ref_init_spade", os.path.join(REF, "nerf", "_init_spade.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    torch.set_num_threads(8)
    m = ref_module()
    gold = os.path.join(REPO, "tests", "golden")
    keys = {}
    for kind, cls in (("generator", m.Generator), ("generator_audio", m.Generator_audio)):
        keys[kind] = {k: list(v.shape) for k, v in cls().state_dict().items()}
    with open(os.path.join(gold, "spade_keys.json"), "w") as f:
        json.dump(keys, f, indent=0)
    import spade_fixtures as SF
    from oracle import spade_oracle as SO

    out = {}
    sd = SF.make_state_dict("generator", seed=0)
    G = m.Generator().eval()
    G.load_state_dict(sd, strict=True)
    for tag, (H, W) in (("64", (64, 64)), ("96x128", (96, 128))):
        inp = SF.make_inputs(H, W, seed=H)
        with torch.no_grad():
            want = G(inp["i_src"], inp["i_raw"])
            got, inter = SO.generator(sd, inp["i_src"], inp["i_raw"], return_intermediates=True)
        err = float((want - got).abs().max())
        print(f"Generator {H}x{W}: oracle vs reference max-abs {err:.3e}; output range [{float(want.min()):.3f}, {float(want.max()):.3f}]")
        assert err == 0.0, "oracle restatement differs from the reference"
        out[f"i_src_{tag}"], out[f"i_raw_{tag}"] = inp["i_src"].numpy(), inp["i_raw"].numpy()
        out[f"ref_out_{tag}"] = want.numpy()
        if tag == "64":
            for k in ("layer2", "layer4", "layer5", "layer6"):
                out[f"ref_{k}_{tag}"] = inter[k].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(gold, "spade_gen.npz"), **out)

    sd = SF.make_state_dict("generator_audio", seed=1)
    Ga = m.Generator_audio().eval()
    Ga.load_state_dict(sd, strict=True)
    inp = SF.make_inputs(64, 64, seed=7)
    with torch.no_grad():
        want = Ga(inp["i_src"], inp["i_raw"], inp["audio"])
        got = SO.generator_audio(sd, inp["i_src"], inp["i_raw"], inp["audio"])
    err = float((want - got).abs().max())
    print(f"Generator_audio 64x64: oracle vs reference max-abs {err:.3e}")
    assert err == 0.0
    np.savez_compressed(os.path.join(gold, "spade_audio.npz"), i_src=inp["i_src"].numpy(), i_raw=inp["i_raw"].numpy(),
                        audio=inp["audio"].numpy(), ref_out=want.numpy())
    print("golden fixtures written")


if __name__ == "__main__":
    main()
