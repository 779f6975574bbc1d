"""CUDA-graph execution of the DDIM loops (fatezero_b200/graphs.py): a configuration seen for the second time is captured per step and
replayed; results must equal the eager run BIT FOR BIT (same kernels, same launch order, deterministic reductions), also when the
prompt-dependent table content changes between replays, and the Python-visible controller state must be the eager one."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from _helpers import build_product, run_product_case  # noqa: E402
from oracle.cases import CASES  # noqa: E402


def _maps(store):
    return [(k, i, t.float().cpu().clone()) for d in store.attention_store_all_step for k, v in d.items() for i, t in enumerate(v)]


@pytest.mark.parametrize("name", ["mini_refine", "mini_replace_blend"])
def test_graph_replay_equals_eager(name, report):
    case = CASES[name]
    pipe = build_product(case["unet"], case["model_config"])
    assert pipe.graph_mode == "auto"
    eager = run_product_case(case, pipe=pipe)                      # first occurrence: eager
    assert not pipe._plans
    maps_eager = _maps(pipe.store_controller)
    acc_eager = {k: [a.float().cpu().clone() for a in v] for k, v in pipe.last_edit_controller.attention_store.items()}
    captured = run_product_case(case, pipe=pipe)                   # second: capture + replay
    assert len(pipe._plans) == 2, list(pipe._plans)
    replayed = run_product_case(case, pipe=pipe)                   # third: replay only
    for tag, got in (("captured", captured), ("replayed", replayed)):
        assert torch.equal(got["inv_latents"], eager["inv_latents"]), tag
        assert torch.equal(got["edit_latents"], eager["edit_latents"]), tag
    # Python-visible state after a replay == eager state
    store = pipe.store_controller
    assert store.cur_step == case["steps"] and len(store.latents_store) == case["steps"]
    for (k, i, a), (k2, i2, b) in zip(maps_eager, _maps(store)):
        assert (k, i) == (k2, i2) and torch.equal(a, b), (k, i)
    ctrl = pipe.last_edit_controller
    assert ctrl.cur_step == case["steps"]
    for k, lst in acc_eager.items():
        for a, b in zip(lst, ctrl.attention_store[k]):
            assert torch.equal(a, b.float().cpu()), k
    if eager["result"]["mask_list"]:
        assert len(replayed["result"]["mask_list"]) == len(eager["result"]["mask_list"])
        for a, b in zip(eager["result"]["mask_list"], replayed["result"]["mask_list"]):
            assert torch.equal(a.cpu(), b.cpu())
    report[f"graphs_{name}"] = dict(plans=len(pipe._plans), graphs=[len(p.steps) for p in pipe._plans.values()],
                                    kernels_per_step=[p.steps.launches[0] for p in pipe._plans.values()])


def test_graph_replay_with_new_prompt(report):
    """Same structure, different target prompt: the captured plan is reused, only the table / text-embedding CONTENT is refreshed."""
    case = CASES["mini_refine"]
    other = copy.deepcopy(case)
    other["target"] = "oil painting of " + case["source"]
    other["p2p"]["eq_params"] = {"words": ["oil"], "values": [4.0]}
    pipe = build_product(case["unet"], case["model_config"])
    run_product_case(case, pipe=pipe)
    run_product_case(case, pipe=pipe)          # captured with the first prompt
    n_plans = len(pipe._plans)
    got = run_product_case(other, pipe=pipe)   # replayed with the second prompt's tables
    assert len(pipe._plans) == n_plans
    ref_pipe = build_product(case["unet"], case["model_config"])
    ref_pipe.graph_mode = "off"
    ref = run_product_case(other, pipe=ref_pipe)
    assert torch.equal(got["inv_latents"], ref["inv_latents"])
    assert torch.equal(got["edit_latents"], ref["edit_latents"])
    assert not torch.equal(got["edit_latents"][-1], run_product_case(case, pipe=ref_pipe)["edit_latents"][-1])
