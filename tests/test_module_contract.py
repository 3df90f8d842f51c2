"""The module_base-shaped host layer (chm_b200/module.py) — what CHM's core sees (CPU)."""
import numpy as np
import pytest

from chm_b200.module import MISSING, Domain, PBSM3D, module_error


def test_depends_provides_match_reference():
    m = PBSM3D({})
    # PBSM3D.cpp:105-110 + :137-140
    assert m.get_depends() == ["U_2m_above_srf", "vw_dir", "swe", "t", "rh", "U_R", "fetch"]
    # PBSM3D.cpp:112-114,144,194-202
    assert m.get_provides() == ["pbsm_more_than_avail", "global_cell_id", "blowingsnow_probability", "Qsubl",
                                "Qsubl_mass", "sum_subl", "drift_mass", "Qsusp", "Qsalt", "sum_drift"]
    assert PBSM3D({"use_tanh_fetch": "false"}).get_depends()[-1] == "p_snow_hours"
    assert PBSM3D.parallel == "domain"


def test_config_errors():
    with pytest.raises(module_error, match="Cannot specify both"):
        PBSM3D({"use_exp_fetch": "true", "use_tanh_fetch": "true"})
    with pytest.raises(module_error, match="unknown config key"):
        PBSM3D({"nlayers": 3})


def test_config_strings_are_coerced_like_ptree():
    m = PBSM3D({"nLayer": "10", "smooth_coeff": "6500", "do_fixed_settling": "true", "settling_velocity": "0.5"})
    assert m.cfg == {"nLayer": 10, "smooth_coeff": 6500, "do_fixed_settling": True, "settling_velocity": 0.5}


def test_domain_store_defaults_to_missing(granger):
    d = Domain(granger, dt=3600)
    d.init_face_data(["Qsusp"])
    assert (d["Qsusp"] == MISSING).all() and d.size_faces() == 985
    with pytest.raises(module_error):
        d["nope"]
    with pytest.raises(module_error, match="before init"):
        PBSM3D({}).run(d)


def test_provider_mirrors_declare_the_reference_contract():
    """scale_wind_vert.cpp:27-45 and fetchr.cpp:27-49: depends / optional / provides and config keys (no GPU needed)."""
    from chm_b200 import module
    sw = module.scale_wind_vert()
    assert sw.get_depends() == ["U_R"] and sw.get_optionals() == ["snowdepthavg"] and sw.get_provides() == ["U_2m_above_srf"]
    assert sw.parallel == "domain" and module.scale_wind_vert(point_mode=True).parallel == "data"
    assert module.scale_wind_vert({"ignore_canopy": "true"}).wind_cfg_kw["ignore_canopy"] == 1
    fe = module.fetchr()
    assert fe.get_depends() == ["vw_dir"] and fe.get_provides() == ["fetch"] and fe.parallel == "data"
    assert fe.wind_cfg_kw == dict(fetch_steps=10, fetch_max_distance=1000.0, fetch_I=0.06, fetch_incl_veg=1)
    assert module.fetchr({"steps": "7", "max_distance": "650", "I": "0.03", "incl_veg": "false"}).wind_cfg_kw == \
        dict(fetch_steps=7, fetch_max_distance=650.0, fetch_I=0.03, fetch_incl_veg=0)
    import pytest
    with pytest.raises(module.module_error):
        module.scale_wind_vert({"ignore_canopi": True})
    with pytest.raises(module.module_error):  # run before the PBSM3D handle exists
        module.fetchr().run(module.Domain.__new__(module.Domain), module.PBSM3D({"nLayer": 5}))


def test_snow_slide_mirror_declares_the_reference_contract():
    """snow_slide.cpp:27-50 (ctor) against the lists read from the compiled reference (tests/golden/golden_slide.npz)."""
    import os
    from chm_b200 import module
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "golden_slide.npz"))
    m = module.snow_slide()
    assert m.get_depends() == list(g["depends"]) and m.get_provides() == list(g["provides"])
    assert m.parallel == "domain"
    assert m.slide_cfg_kw == dict(avalache_mult=3178.4, avalache_pow=-1.998, use_vertical_snow=1)
    assert module.snow_slide({"avalache_mult": "900", "use_vertical_snow": "false"}).slide_cfg_kw["avalache_mult"] == 900.0
    with pytest.raises(module.module_error):
        module.snow_slide({"avalanche_mult": 900})          # the reference's key is spelt "avalache"
    with pytest.raises(module.module_error):
        module.snow_slide().run(module.Domain.__new__(module.Domain), module.PBSM3D({"nLayer": 5}))


def test_smooth_coeff_is_read_as_an_int_like_the_reference(monkeypatch):
    """PBSM3D.cpp:235 cfg.get("smooth_coeff", 820): an int read; ptree returns the default for text that is not an int."""
    from chm_b200 import capi, module
    seen = {}

    class FakeHandle:
        def __init__(self, cc, *a, **k):
            seen["smooth_coeff"] = cc.smooth_coeff

    monkeypatch.setattr(capi, "Handle", FakeHandle)
    import conftest
    for text, want in (("6500", 6500.0), ("820.7", 820.0), (900.0, 900.0)):
        m = module.PBSM3D({"smooth_coeff": text})
        m.init(module.Domain(conftest.load_mesh("granger1m")))
        assert seen["smooth_coeff"] == want, text
