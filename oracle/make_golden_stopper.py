"""Test infrastructure, not product code.  Generates tests/golden/early_stopper_trace.pt: the reference's own
``AnchoredEarlyStopper`` (delta_experiment/scripts/early_stopping.py:72-317, imported from /root/reference through
oracle/ref_bridge.py) driven with scripted anchor losses, so that the decisions of the state machine -- when it checks,
what it reports, when it stops, which snapshot it restores -- are pinned independently of any DiT arithmetic.

Run here (needs /root/reference):  python oracle/make_golden_stopper.py"""
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_bridge  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

# (name, constructor kwargs, anchor loss at setup, then one scripted loss per CHECK in order)
CASES = [
    ("patience3_every5", dict(check_every=5, patience=3), [1.00, 0.90, 0.95, 0.80, 0.85, 0.86, 0.87, 0.70]),
    ("patience1_every1", dict(check_every=1, patience=1), [1.00, 0.90, 0.80, 0.80, 0.70]),
    ("patience2_every2_never_improves", dict(check_every=2, patience=2), [1.00, 1.10, 1.20, 0.50]),
    ("first_rise_every1", dict(check_every=1, patience=3, strategy="first_rise"), [1.00, 0.90, 0.80, 0.81, 0.10]),
    ("first_rise_every3_immediate", dict(check_every=3, patience=5, strategy="first_rise"), [1.00, 1.00, 0.50]),
    ("patience3_every4_monotone", dict(check_every=4, patience=3), [1.00, 0.90, 0.80, 0.70, 0.60, 0.50]),
]
N_STEPS = 40


def drive(stopper_cls, kwargs, losses):
    """The loop shape of run_lora_tta.py:473-541: setup, then step(step + 1) after every optimizer step, snapshots are the
    step index at which they were taken."""
    st = stopper_cls(**kwargs)
    seq = iter(losses)
    st._compute_anchor_loss = lambda: next(seq)          # scripted anchor losses
    current = {"step": 0}
    save_fn = lambda: {"snapshot_of_step": current["step"]}   # noqa: E731
    model = torch.nn.Linear(1, 1)
    z = torch.zeros(1, 1, 1, 1, 1)
    st.setup(model, z, z, z, torch.ones(1, 1), device="cpu", dtype=torch.float32, video_id="trace", save_fn=save_fn)
    trace = []
    for step in range(N_STEPS):
        current["step"] = step + 1
        try:
            stop, info = st.step(step + 1, save_fn=save_fn)
        except StopIteration:       # script exhausted: the loop would have had to stop before
            trace.append("script exhausted")
            break
        trace.append((step + 1, bool(stop), dict(info)))
        if stop:
            break
    restored = {}
    st.restore(restore_fn=lambda s: restored.update(s))
    return {"trace": trace, "state": st.state, "restored": restored}


def main():
    es = ref_bridge.load("early_stopping")
    out = {name: {"kwargs": kw, "losses": ls, **drive(es.AnchoredEarlyStopper, kw, ls)} for name, kw, ls in CASES}
    for k, v in out.items():
        print(k, "->", v["trace"][-1], v["restored"])
    torch.save(out, GOLDEN / "early_stopper_trace.pt")


if __name__ == "__main__":
    main()
