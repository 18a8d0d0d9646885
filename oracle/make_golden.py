"""Generate the committed golden fixtures under tests/golden/ by running the
UNMODIFIED reference (through oracle/ref_harness.py) in the build container.

TEST INFRASTRUCTURE ONLY.  Usage (container, /root/reference present):

    python -m oracle.make_golden            # regenerate everything (a few minutes)
    python -m oracle.make_golden tick_1000  # only fixtures whose name starts with ...

Fixture kinds (all compressed .npz; see tests/golden/README.md):
  tend_*.npz  state just before the reference's calculate_criminal_tendency +
              calculate_crime_multiplier, and their results
              (model.py:1144-1186).
  tick_*.npz  state just before reset_oc_embeddedness + commit_crimes
              (model.py:272-273), the raw uniform draws the reference consumed
              (R1-R6), per-crime records (initiator, k, rings, candidate keys,
              group), every oc_embeddedness evaluation (accumulated and
              fresh-meta-graph values), arrests, and the state afterwards.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import ref_harness as H

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

PARAM_KEYS = ["nat_propensity_m", "nat_propensity_sigma", "nat_propensity_threshold",
              "number_crimes_yearly_per10k", "number_arrests_per_year", "punishment_length",
              "threshold_use_facilitators", "facilitator_repression_multiplier", "good_guy_threshold",
              "max_accomplice_radius", "oc_embeddedness_radius", "facilitator_repression",
              "oc_boss_repression", "ticks_between_intervention", "intervention_start", "intervention_end",
              "this_is_a_big_crime", "ticks_per_year"]


def _params(model) -> dict:
    return {"param_" + k: np.float64(getattr(model, k)) for k in PARAM_KEYS}


def _pack_state(prefix: str, st: dict) -> dict:
    return {prefix + k: v for k, v in st.items()}


def pack_tick(model, h: dict) -> dict:
    pre, post, rec = h["pre"], h["post"], h["rec"]
    slot_of = {int(u): j for j, u in enumerate(pre["unique_id"])}
    assert np.array_equal(pre["unique_id"], post["unique_id"])
    out = {}
    out.update(_pack_state("pre_", pre))
    # post: only what the hot path can change
    for k in ["oc_member", "prisoner", "num_crimes_committed", "num_crimes_committed_this_tick", "job_level",
              "sentence_countdown", "new_recruit", "has_job", "has_school", "arrest_weight",
              "rowptr_6", "col_6", "co_off", "rowptr_7", "col_7", "rowptr_8", "col_8"]:
        out["post_" + k] = post[k]
    out.update(_params(model))
    out["num_co_offenders_dist"] = np.array(model.num_co_offenders_dist, dtype=np.float64)
    out["tick"] = np.int32(h["pre_scalars"]["tick"])
    out["crime_multiplier"] = np.float64(h["pre_scalars"]["crime_multiplier"])
    out["n_alive"] = np.int32(h["n_alive"])
    out["intervention_on"] = np.int32(h["intervention_on"])
    for k in ["number_crimes", "crime_size_fails", "facilitator_crimes", "facilitator_fails",
              "big_crime_from_small_fish", "number_offspring_recruited_this_tick",
              "number_protected_recruited_this_tick", "people_jailed"]:
        out["d_" + k] = np.int32(h["post_scalars"][k] - h["pre_scalars"][k])
    rp = H.replay_from_record(rec, slot_of)
    for k, v in rp.items():
        out["replay_" + k] = np.asarray(v)
    crimes = rec["crimes"]
    out["crime_initiator"] = np.array([slot_of[c["initiator"]] for c in crimes], dtype=np.int32)
    out["crime_k"] = np.array([c["k"] for c in crimes], dtype=np.int32)
    out["crime_oc"] = np.array([c["oc"] for c in crimes], dtype=np.uint8)
    out["crime_d_fails"] = np.array([c["d_fails"] for c in crimes], dtype=np.int32).reshape(-1, 3)
    goff, gmem = [0], []
    koff, kslot, kval = [0], [], []
    roff, rd, rmoff, rmem = [0], [], [0], []
    for c in crimes:
        gmem.extend(slot_of[u] for u in c["group"])  # reference order: list(set) + [initiator]
        goff.append(len(gmem))
        for u, v in c["keys"].items():
            kslot.append(slot_of[u]); kval.append(v)
        koff.append(len(kslot))
        for d, members in c["rings"]:
            rd.append(d)
            rmem.extend(slot_of[u] for u in members)
            rmoff.append(len(rmem))
        roff.append(len(rd))
    out["group_off"], out["group_mem"] = np.array(goff, dtype=np.int32), np.array(gmem, dtype=np.int32)
    out["key_off"], out["key_slot"] = np.array(koff, dtype=np.int32), np.array(kslot, dtype=np.int32)
    out["key_val"] = np.array(kval, dtype=np.float64)
    out["ring_off"], out["ring_d"] = np.array(roff, dtype=np.int32), np.array(rd, dtype=np.int32)
    out["ring_mem_off"], out["ring_mem"] = np.array(rmoff, dtype=np.int32), np.array(rmem, dtype=np.int32)
    out["emb_slot"] = np.array([slot_of[u] for u, _, _ in rec["emb"]], dtype=np.int32)
    out["emb_accumulated"] = np.array([v for _, v, _ in rec["emb"]], dtype=np.float64)
    out["emb_fresh"] = np.array([v for _, _, v in rec["emb"]], dtype=np.float64)
    out["arrested"] = np.array([slot_of[u] for u in rec["arrested"]], dtype=np.int32)
    # meta-edge direction winners on pairs with asymmetric multiplicities: (evaluation index, slot a, slot b, multiplicity)
    out["emb_over"] = np.array([(j, slot_of[a], slot_of[b], m) for j, a, b, m in rec.get("emb_over", [])],
                               dtype=np.int32).reshape(-1, 4)
    return out


def pack_report(model, h: dict) -> dict:
    """State at the end of a reference step() and the reporter values it collected from it (model.py:286-287)."""
    out = _pack_state("pre_", h["end"])
    out.update(_params(model))
    for k, v in h["end_reporters"].items():
        out["rep_" + k] = np.float64(v)
    out["tick"] = np.int32(model.tick)
    return out


def pack_tendency(model) -> dict:
    """State right before the yearly tendency pass and its results."""
    pre = H.export_state(model)
    model.calculate_criminal_tendency()
    model.calculate_crime_multiplier()
    out = _pack_state("pre_", pre)
    out.update(_params(model))
    post = H.export_state(model)
    assert np.array_equal(pre["unique_id"], post["unique_id"])
    out["post_criminal_tendency"] = post["criminal_tendency"]
    out["post_crime_multiplier"] = np.float64(model.crime_multiplier)
    out["tick"] = np.int32(model.tick)
    return out


def save(name: str, d: dict):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def run_case(prefix, n_agents, seed, ticks_to_record, overrides=None, tend_ticks=(), pre_setup=None, post_setup=None,
             only=None, rep_ticks=()):
    names = ["tick_%s_t%d" % (prefix, t) for t in ticks_to_record] + ["tend_%s_t%d" % (prefix, t) for t in tend_ticks]
    if only and not any(n.startswith(only) for n in names):
        return
    ref_model, _, _ = H.load_reference()
    model = ref_model.ProtonOC(seed=seed)
    if overrides:
        model.override_dict(overrides)
    if pre_setup:
        pre_setup(model)
    model.setup(n_agents)
    if post_setup:
        post_setup(model)
    last = max(list(ticks_to_record) + list(tend_ticks))
    for _ in range(last):
        tick = model.tick
        if tick in tend_ticks:
            # the tendency pass is idempotent given the state, so record it on a scratch call
            # just before step() runs it (model.py:248-250); ages must be current first
            for a in model.schedule.agents:
                a.calculate_age()
            d = pack_tendency(model)
            save("tend_%s_t%d" % (prefix, tick), d)
        if tick in ticks_to_record:
            h = H.step_with_recording(model)
            save("tick_%s_t%d" % (prefix, tick), pack_tick(model, h))
            if tick in rep_ticks:
                save("rep_%s_t%d" % (prefix, tick), pack_report(model, h))
        else:
            model.step()


def main(argv):
    """argv[1]: only fixtures whose name starts with this prefix; `case:<prefix>` selects one run_case by its prefix
    (so that the cases can be generated by parallel processes)."""
    only = argv[1] if len(argv) > 1 else None
    case = None
    if only and only.startswith("case:"):
        case, only = only[5:], None

    def rc(prefix, *a, **kw):
        if case is None or prefix == case:
            run_case(prefix, *a, only=only, **kw)

    # config 1 of BASELINE.json: 100 agents, seed 42
    rc("n100_s42", 100, 42, [1, 12, 60], tend_ticks=[1, 12], rep_ticks=[12, 60])
    # 1000 agents: early ticks, yearly tick, and a later tick with OC growth
    rc("n1000_s42", 1000, 42, [1, 5, 12, 13, 120], tend_ticks=[1, 12, 120], rep_ticks=[13, 120])
    # 3000 agents (config 2): embeddedness evaluations appear
    rc("n3000_s42", 3000, 42, [1, 24, 30], tend_ticks=[1], rep_ticks=[24])
    # facilitator repression active every tick from 13 (model.py:1326-1336)
    rc("n1000_s7_facilitators", 1000, 7, [13, 20], overrides={"intervention": "facilitators"})
    # OC-boss repression (model.py:1293-1302), more OC members so the regime is exercised
    rc("n1000_s7_disruptive_strong", 1000, 7, [14, 25],
       overrides={"intervention": "disruptive-strong", "num_oc_persons": 200, "num_oc_families": 30}, rep_ticks=[25])

    # the reference's own test_big_crimes_from_small_fish setting (test_protonoc.py:103-114)
    def big_post(model):
        model.num_co_offenders_dist = [[5, 0.5], [6, 0.5], [10, 0.5]]
    rc("n500_s3_radius6", 500, 3, [2, 7], overrides={"max_accomplice_radius": 6}, post_setup=big_post)
    # many OC members at 1000 agents so OC-initiated crimes and embeddedness dominate
    rc("n1000_s11_oc200", 1000, 11, [3, 15, 40], overrides={"num_oc_persons": 300, "num_oc_families": 8})
    # late in a run: twenty simulated years of the reference's own demography, arrests and co-offending behind it
    # (one-sided clears, a dense criminal layer, negative tendencies)
    rc("n1000_s5_late", 1000, 5, [240, 252], tend_ticks=[240], rep_ticks=[252])
    # ... and the same with a large OC population, where arrests, recruitment and embeddedness have been busy for ten years
    rc("n1000_s13_oc150_late", 1000, 13, [120, 121], tend_ticks=[120],
       overrides={"num_oc_persons": 150, "num_oc_families": 8})

    # ---- round 2: the OC-initiated branch (the crimes that recruit, entities.py:465-467) and full-size states
    # group sizes skewed to large groups so that most crimes search for accomplices; 300 of 1,000 agents in OC families
    def skew_post(model):
        model.num_co_offenders_dist = [[1, 0.15], [2, 0.3], [3, 0.25], [4, 0.15], [5, 0.1], [6, 0.05]]
    rc("n1000_s17_oc300_big", 1000, 17, list(range(2, 14)), overrides={"num_oc_persons": 300, "num_oc_families": 8},
       post_setup=skew_post)
    rc("n3000_s19_oc300_big", 3000, 19, [2, 3, 4, 5], overrides={"num_oc_persons": 300, "num_oc_families": 24},
       post_setup=skew_post)
    # BASELINE config 3/5 size: 10,000 agents, baseline parameters, a quarter into the second year (= the bench workload
    # state_n10000_s42_t13) and the tick after it; tendency at setup and at the first yearly pass
    rc("n10000_s42", 10000, 42, [13, 14], tend_ticks=[1, 12], rep_ticks=[13])
    # samples/sample_json.json run "0": 80 OC members in 16 families, facilitator repression
    rc("n10000_s42_sample0", 10000, 42, [1, 2, 13],
       overrides={"intervention": "facilitators", "num_oc_persons": 80, "number_arrests_per_year": 20, "num_oc_families": 16})


if __name__ == "__main__":
    main(sys.argv)
