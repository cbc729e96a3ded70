"""Export path (reference scripts/runtime/compile.py:160-210): `torch.jit.trace(model, (inputs,), strict=False)` on the
mirror records the `creste::` dispatcher ops (creste_public_b200/torch_ops.py); the traced module reproduces the eager
forward on NEW inputs, survives save / load, and its graph holds no Python call-backs."""
import io

import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu
H, W = 64, 96


@pytest.mark.parametrize("mode", ["fp32", "3xfp16"])
def test_traced_model_equals_eager(cuda, mode, tmp_path):
    import creste_public_b200 as cb
    cb.set_precision(mode)
    try:
        model = cb.build_maxentirl(image_size=(H, W)).eval()
        model.load_state_dict(synth.seeded_state_dict(model.state_dict(), 0, "soft"))
        model = model.cuda()
        a, p2p = synth.net_inputs(H, W, 1, seed=0)
        b, _ = synth.net_inputs(H, W, 1, seed=1)
        inputs = (a.cuda(), p2p.cuda())
        with torch.no_grad():
            traced = torch.jit.trace(model, (inputs,), strict=False)
        graph = str(traced.inlined_graph)
        for op in ("creste::conv2d", "creste::dwconv_bn_swish", "creste::se_gate", "creste::splat_soft",
                   "creste::frustum_to_bev", "creste::depth_expectation", "creste::proj_head", "creste::maxpool2_concat"):
            assert op in graph, op
        assert "PythonOp" not in graph
        with torch.no_grad():
            eager = model((b.cuda(), p2p.cuda()))
            got = traced((b.cuda(), p2p.cuda()))
        assert set(got.keys()) == set(eager.keys())
        # fp32: the traced ops ARE the eager ops.  3xfp16: the eager path lets conv epilogues write the next conv's
        # fp16 hi / lo operand with a scale from an a-priori bound, the traced (pure-function) ops derive it from the
        # carried maximum -- two power-of-two scales, identical products except for elements that are fp16-subnormal
        # under the looser one: last-bit differences, bounded here at 2e-6 of the tensor maximum; the integers agree
        for k in ("depth_preds_feats", "depth_preds_logits", "dino_pe_feats"):
            if mode == "fp32":
                assert torch.equal(got[k], eager[k]), k
            else:
                assert float((got[k] - eager[k]).abs().max()) <= 2e-6 * float(eager[k].abs().max()), k
        assert torch.equal(got["depth_preds_bins"], eager["depth_preds_bins"])
        if mode == "fp32":
            assert torch.equal(got["bev_coords"], eager["bev_coords"])
        else:       # float voxel coordinates: 10 voxels per metre of a depth that differs in its last bits
            assert float((got["bev_coords"] - eager["bev_coords"]).abs().max()) <= 1e-3
        for k in ("bev_features", "inpainting_sam_preds", "elevation_features", "input_view", "traversability_preds",
                  "traversability_preds_full"):
            assert float((got[k] - eager[k]).abs().max()) <= 1e-4 * max(1.0, float(eager[k].abs().max())), k   # splat atomics
        # save / load round trip (the C++ runtime loads this file; here the Python-registered ops serve it)
        path = str(tmp_path / "traversability_model_trace.pt")
        traced.save(path)
        loaded = torch.jit.load(path)
        with torch.no_grad():
            again = loaded((b.cuda(), p2p.cuda()))
        assert torch.equal(again["depth_preds_feats"], got["depth_preds_feats"])
        assert float((again["traversability_preds"] - eager["traversability_preds"]).abs().max()) <= 1e-4
    finally:
        cb.set_precision("fp32")


def test_trace_batch_generalises(cuda):
    """The trace is shape-specialised only where the reference's is (python ints of the input shape): replaying on the
    traced shape with other values is exact; the eager path is untouched by having traced."""
    import creste_public_b200 as cb
    cb.set_precision("fp32")
    model = cb.build_terrainnet(image_size=(H, W)).eval()
    model.load_state_dict(synth.seeded_state_dict(model.state_dict(), 0, "peaky"))
    model = model.cuda()
    a, p2p = synth.net_inputs(H, W, 2, seed=3)
    with torch.no_grad():
        before = model((a.cuda(), p2p.cuda()))["elevation_preds"].clone()
        traced = torch.jit.trace(model, ((a.cuda(), p2p.cuda()),), strict=False)
        after = model((a.cuda(), p2p.cuda()))["elevation_preds"]
        t = traced((a.cuda(), p2p.cuda()))["elevation_preds"]
    assert float((before - after).abs().max()) <= 1e-4 and float((t - after).abs().max()) <= 1e-4


RUNTIME = r'''
import sys, torch
torch.ops.load_library(sys.argv[1])                      # C++ registration of creste:: over libcreste_b200.so
assert "creste_public_b200" not in sys.modules
m = torch.jit.load(sys.argv[2])
io = torch.load(sys.argv[3])
with torch.no_grad():
    out = m((io["rgbd"].cuda(), io["p2p"].cuda()))
assert "creste_public_b200" not in sys.modules            # no product Python was needed to run the traced model
assert torch.equal(out["depth_preds_feats"].cpu(), io["feats"]) and torch.equal(out["depth_preds_bins"].cpu(), io["bins"])
err = float((out["traversability_preds"].cpu() - io["costmap"]).abs().max())
assert err <= 1e-4, err
print("OK", err)
'''


def test_traced_model_runs_on_the_cpp_registered_ops(cuda, tmp_path):
    """The TorchScript file the reference's compile.py would save runs in a process that never imports the Python
    package: the `creste::` ops come from csrc_torch/libcreste_torch_ops.so (C++ TORCH_LIBRARY over the C ABI) -- the
    library a libtorch runtime links."""
    import os
    import subprocess
    import sys
    import creste_public_b200 as cb
    from creste_public_b200.csrc_torch import build as tb_build
    lib = tb_build.build()
    cb.set_precision("3xfp16")
    try:
        model = cb.build_maxentirl(image_size=(H, W)).eval()
        model.load_state_dict(synth.seeded_state_dict(model.state_dict(), 0, "soft"))
        model = model.cuda()
        a, p2p = synth.net_inputs(H, W, 1, seed=0)
        b, _ = synth.net_inputs(H, W, 1, seed=2)
        with torch.no_grad():
            traced = torch.jit.trace(model, ((a.cuda(), p2p.cuda()),), strict=False)
            want = model((b.cuda(), p2p.cuda()))
    finally:
        cb.set_precision("fp32")
    path, io = str(tmp_path / "model.pt"), str(tmp_path / "io.pt")
    traced.save(path)
    torch.save({"rgbd": b, "p2p": p2p, "feats": want["depth_preds_feats"].cpu(), "bins": want["depth_preds_bins"].cpu(),
                "costmap": want["traversability_preds"].cpu()}, io)
    script = tmp_path / "runtime.py"
    script.write_text(RUNTIME)
    res = subprocess.run([sys.executable, str(script), lib, path, io], capture_output=True, text=True, timeout=600,
                         cwd=str(tmp_path), env={k: v for k, v in os.environ.items() if k != "PYTHONPATH"})
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
