"""Host-side parameter block of the crime path: turns ProtonOC-style attributes and tables into poc_params.

CDFs are built the way the reference's sampling primitive builds them (extra.weighted_n_of, extra.py:114-128,
then numpy Generator.choice: cumsum(p); cdf /= cdf[-1]) so that searchsorted(cdf, U, 'right') on the device
returns what the reference returns for the same uniform U."""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import numpy as np

from . import tables
from ._cabi import MAX_CELLS, MAX_TAB, Demography, Labour, Params

# the model attributes the hot path reads (model.py:108-134, 82-83) with the reference defaults
HOT_PATH_DEFAULTS = dict(
    nat_propensity_m=1.0, nat_propensity_sigma=0.25, nat_propensity_threshold=1.0,
    number_crimes_yearly_per10k=2000, number_arrests_per_year=30, punishment_length=1,
    threshold_use_facilitators=4, facilitator_repression_multiplier=2.0, good_guy_threshold=0.6,
    max_accomplice_radius=2, oc_embeddedness_radius=2, facilitator_repression=False,
    oc_boss_repression=False, ticks_between_intervention=12, intervention_start=13,
    intervention_end=999, this_is_a_big_crime=3, ticks_per_year=12)


def choice_cdf(weights: Iterable[float]) -> np.ndarray:
    p = [float(x) for x in weights]
    sump = sum(p)                       # Python left-to-right sum, extra.py:120
    p = np.array([i / sump for i in p], dtype=np.float64)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return cdf


def propensity_threshold(m: float, sigma: float, thr: float) -> float:
    """The expression at entities.py:350-353, verbatim (Q4)."""
    return float(np.exp(m - sigma ** 2 / 2) + thr * np.sqrt(np.exp(sigma) ** 2 - 1) * np.exp(m + sigma ** 2 / 2))


def make_params(attrs: Optional[Dict] = None, c_range=None, co_offenders=None, conviction=None) -> Params:
    """attrs: any mapping / object attributes named like the reference's model attributes."""
    d = dict(HOT_PATH_DEFAULTS)
    if attrs is not None:
        get = attrs.get if isinstance(attrs, dict) else (lambda k, default=None: getattr(attrs, k, default))
        for k in d:
            v = get(k, None)
            if v is not None:
                d[k] = v
    c_range = c_range if c_range is not None else tables.C_RANGE_BY_AGE_AND_SEX
    co_offenders = co_offenders if co_offenders is not None else tables.NUM_CO_OFFENDERS_DIST
    conviction = conviction if conviction is not None else tables.CONVICTION_LENGTH
    if len(c_range) > MAX_CELLS or len(co_offenders) > MAX_TAB or len(conviction) > MAX_TAB:
        raise ValueError("table too large for poc_params")
    p = Params()
    p.n_cells = len(c_range)
    for i, (male, a0, a1, c) in enumerate(c_range):
        p.cell_male[i], p.cell_age_from[i], p.cell_age_to[i], p.cell_c[i] = int(male), int(a0), int(a1), float(c)
    p.n_sizes = len(co_offenders)
    cdf = choice_cdf([r[-1] for r in co_offenders])
    for i, r in enumerate(co_offenders):
        p.size_n[i], p.size_cdf[i] = int(r[0]), float(cdf[i])
    p.n_sent = len(conviction)
    cf, cm = choice_cdf([r[1] for r in conviction]), choice_cdf([r[2] for r in conviction])
    for i, r in enumerate(conviction):
        p.sent_months[i], p.sent_cdf_f[i], p.sent_cdf_m[i] = float(r[0]), float(cf[i]), float(cm[i])
    p.propensity_threshold = propensity_threshold(d["nat_propensity_m"], d["nat_propensity_sigma"],
                                                  d["nat_propensity_threshold"])
    p.crimes_yearly_per10k = float(d["number_crimes_yearly_per10k"])
    p.arrests_per_year = float(d["number_arrests_per_year"])
    p.punishment_length = float(d["punishment_length"])
    p.threshold_use_facilitators = float(d["threshold_use_facilitators"])
    p.facilitator_repression_multiplier = float(d["facilitator_repression_multiplier"])
    p.good_guy_threshold = float(d["good_guy_threshold"])
    for k in ["max_accomplice_radius", "oc_embeddedness_radius", "facilitator_repression", "oc_boss_repression",
              "ticks_between_intervention", "intervention_start", "intervention_end", "this_is_a_big_crime",
              "ticks_per_year"]:
        setattr(p, k, int(d[k]))
    return p


def make_demography(nat_propensity_m: float = 1.0, nat_propensity_sigma: float = 0.25, mortality=None, fertility=None,
                    retirement_age: int = 65, number_weddings_mean: float = None) -> Demography:
    """poc_demography for the phases of SURVEY.md 8(f).  mortality: (female[ages], male[ages]) yearly probabilities
    (initial_mortality_rates.csv, entities.py:615-626), fertility: [3][64] yearly by [min(children, 2)][age]
    (initial_fertility_rates.csv, entities.py:605-613); defaults: the reference's Palermo tables."""
    fem, mal = mortality if mortality is not None else (tables.MORTALITY_FEMALE, tables.MORTALITY_MALE)
    fert = fertility if fertility is not None else tables.FERTILITY
    d = Demography()
    for a in range(128):
        d.mort[0][a] = float(fem[a]) if a < len(fem) else 1.0
        d.mort[1][a] = float(mal[a]) if a < len(mal) else 1.0
    for k in range(3):
        for a in range(64):
            d.fert[k][a] = float(fert[k][a])
    for i, v in enumerate(tables.poisson_cdf(3.0, 32)):
        d.poisson_cdf[i] = v
    for i, v in enumerate(tables.lognormal_quantiles(nat_propensity_m, nat_propensity_sigma, 1024)):
        d.prop_q[i] = v
    d.mort_max_age, d.n_prop = len(fem) - 1, 1024
    d.retirement_age = int(retirement_age)   # model.py:121
    d.weddings_mean = float(tables.NUMBER_WEDDINGS_MEAN if number_weddings_mean is None else number_weddings_mean)   # model.py:184-186
    return d


def fill_labour(lb, education_levels=None, work_status=None, wealth_quintile=None, labour_status=None, range_ages=None):
    """Fill a poc_labour-shaped ctypes struct (the oracle's mirror has the same layout).  Tables as the reference holds them:
    education_levels {level: (enter age, escape age)} (model.py:953-965), work_status {education level: {male: [[status,
    rate], ...]}} (model.py:166-167), wealth_quintile {work status: {male: [[wealth, rate], ...]}} (:168-169), labour_status
    {male: {age: share}} (:180-181), range_ages {male: [ages]} (:182-183); defaults: the reference's Palermo tables."""
    education_levels = education_levels if education_levels is not None else tables.EDUCATION_LEVELS
    work_status = work_status if work_status is not None else tables.WORK_STATUS_BY_EDU_LVL
    wealth_quintile = wealth_quintile if wealth_quintile is not None else tables.WEALTH_QUINTILE_BY_WORK_STATUS
    labour_status = labour_status if labour_status is not None else tables.LABOUR_STATUS_BY_AGE_AND_SEX
    range_ages = range_ages if range_ages is not None else tables.LABOUR_STATUS_RANGE_AGES
    if not education_levels or sorted(education_levels) != list(range(1, len(education_levels) + 1)) or len(education_levels) > 7:
        raise ValueError("education levels must be numbered 1..n, n <= 7")
    lb.n_levels = len(education_levels)
    lb.primary_age = int(education_levels[1][0])
    for lv, row in education_levels.items():
        lb.end_age[lv] = int(row[1])
    for tab, n_, val_, cdf_ in ((work_status, lb.work_n, lb.work_val, lb.work_cdf),
                                (wealth_quintile, lb.wealth_n, lb.wealth_val, lb.wealth_cdf)):
        for key, by_sex in tab.items():
            for male, rows in by_sex.items():
                if not 0 <= int(key) < 8 or len(rows) > 8:
                    raise ValueError("table too large for poc_labour")
                cdf = choice_cdf([r[-1] for r in rows])          # extra.pick_from_pair_list: weighted_one_of over the rates
                n_[int(key)][int(bool(male))] = len(rows)
                for i, r in enumerate(rows):
                    val_[int(key)][int(bool(male))][i] = int(r[0])
                    cdf_[int(key)][int(bool(male))][i] = float(cdf[i])
    for male in (0, 1):
        for a in range(128):
            lb.inactive[male][a] = float(labour_status[bool(male)].get(a, -1.0))   # no row: the reference raises KeyError
            lb.range_age[male][a] = 1 if a in range_ages[bool(male)] else 0
    return lb


def make_labour(**tables_) -> Labour:
    """poc_labour for CrimeEngine.set_labour (graduate_and_enter_jobmarket + update_unemployment_status on the device)."""
    return fill_labour(Labour(), **tables_)
