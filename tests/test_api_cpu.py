"""Host-logic tests of the reference-facing interface (no GPU): argument handling and error behaviour of
mcmcglm() as in R/mcmcglm.R:159-169, model-frame extraction, accessor layout."""
import numpy as np
import pandas as pd
import pytest
import mcmcglm_b200 as mg
from mcmcglm_b200 import _lib


def _dat(n=20):
    rng = np.random.default_rng(0)
    return pd.DataFrame({"Y": rng.standard_normal(n), "X1": rng.standard_normal(n), "X2": rng.integers(0, 2, n)})


def test_reference_guards():
    d = _dat()
    with pytest.raises(ValueError, match="Need more iterations than burnin"):          # R/mcmcglm.R:165
        mg.mcmcglm("Y ~ .", "gaussian", d, w=0.5, n_samples=10, burnin=10)
    with pytest.raises(ValueError, match="A tuning parameter for the `qslice_fun` is missing"):   # :167-169
        mg.mcmcglm("Y ~ .", "gaussian", d)
    with pytest.raises(ValueError, match="should be one of"):                              # match.arg :161
        mg.mcmcglm("Y ~ .", "gaussian", d, w=0.5, linear_predictor_calc="fast")
    with pytest.raises(ValueError, match="should be one of"):                              # match.arg :163
        mg.mcmcglm("Y ~ .", "gaussian", d, w=0.5, sample_method="hmc")


@pytest.mark.parametrize("kw", [
    dict(family="Gamma"), dict(family=mg.binomial(link="cloglog")), dict(family=mg.poisson(link="identity")),
    dict(beta_prior=[mg.dist_normal()] * 9), dict(beta_prior="gamma"), dict(sample_method="normal-normal"),
    dict(qslice_fun=lambda **k: None)])
def test_unsupported_inputs_are_rejected_not_emulated(kw):
    args = dict(formula="Y ~ .", family="gaussian", data=_dat(), w=0.5)
    args.update(kw)
    with pytest.raises(mg.CggError) as e:
        mg.mcmcglm(**args)
    assert e.value.code in (_lib.E_UNSUPPORTED,)


def test_model_frame():
    d = _dat()
    Y, X, names = mg.extract_model_data("Y ~ .", d)
    assert names == ["(Intercept)", "X1", "X2"] and X.shape == (20, 3) and np.all(X[:, 0] == 1)
    assert X.flags.f_contiguous                      # column-major like an R matrix
    _, X2, names2 = mg.extract_model_data("Y ~ X2 + X1 - 1", d)
    assert names2 == ["X2", "X1"] and np.array_equal(X2[:, 1], d["X1"].to_numpy())
    with pytest.raises(ValueError):
        mg.extract_model_data("Y ~ X1:X2", d)
    assert mg.check_family("poisson").link == "log" and mg.check_family(mg.binomial).family == "binomial"


def test_accessors_on_a_synthetic_object():
    # Q1: burnin flag is iteration <= burnin + 1; Q2: quantile() summarises burnin == TRUE rows; Q3: coef over FALSE rows
    it = np.arange(11)
    df = pd.DataFrame({"(Intercept)": it * 1.0, "X1": it * 2.0, "iteration": it, "burnin": it <= 3 + 1})
    x = mg.McmcGlm(beta_samples=df, beta_mean=df.loc[~df["burnin"], ["(Intercept)", "X1"]].mean().to_frame().T, data=None,
                   model_matrix=np.zeros((5, 2)), param_list=None, family=mg.gaussian(), formula="Y ~ .", call="mcmcglm(...)",
                   burnin=3, sample_method="slice_sampling", qslice_fun=mg.slice_stepping_out, tuning={"w": 0.5})
    assert mg.samples(x) is df and mg.coef(x).iloc[0, 0] == np.mean(it[5:])
    q = mg.quantile(x)
    assert list(q.columns) == ["var", "mean", "q_0025", "q_05", "q_0975"]
    assert q.loc[0, "mean"] == 2.0 and q.loc[1, "q_05"] == 4.0           # rows 0..4 only
    assert x.w == 0.5 and "Object of class 'mcmcglm'" in repr(x)
