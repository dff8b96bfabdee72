"""Pins every BxDF kind of the oracle against formulas restated HERE in numpy from the reference's sources — independent of
oracle/ref_render.cpp — plus the integral identities a correct f / pdf / sample_f triple satisfies whatever its formula:

  * f(wo, wi) at random direction pairs equals the numpy restatement (Lambertian, OrenNayar, Microfacet x Blinn with the
    dielectric and the conductor Fresnel term, FresnelBlend is pinned in test_oracle_materials.py);
  * Helmholtz reciprocity f(wo, wi) == f(wi, wo);
  * the directional albedo  rho(wo) = int f cos  is <= 1 for reflectances <= 1 (quadrature);
  * pdf integrates to its closed form over the sphere of directions (1 for the cosine density; 1 - 2^-(e+1)/2 for the
    Blinn half-vector density seen from the normal, where reflections below the horizon are dropped);
  * sample_f draws from pdf: the Monte Carlo estimate  mean(f cos / pdf)  over a stratified (u1, u2) grid reproduces the
    quadrature of  f cos,  and the sampled direction's pdf is the pdf() of that direction;
  * BSDF-level bookkeeping (bsdf.dart:53-198): component choice floor(u * matching), pdf averaged over the matching BxDFs,
    f summed, reflection / transmission gated by the geometric normal; BRDFToBTDF mirrors wi; ScaledBxDF scales f and keeps
    the cosine density;
  * specular lobes: Fresnel energy split of glass, Snell's law, reciprocity of the dielectric term across the interface,
    the conductor term's exact values at normal and grazing incidence.

Reference sources restated: lib/core/reflection/{lambertian,oren_nayar,microfacet,blinn,fresnel_dielectric,fresnel_conductor,
specular_reflection,specular_transmission,brdf_to_btdf,scaled_bxdf,bxdf,bsdf}.dart."""
import math

import numpy as np
import pytest

from dartray_b200 import host
from tests.oracle_lib import Oracle

BSDF_REFLECTION, BSDF_TRANSMISSION, BSDF_DIFFUSE, BSDF_GLOSSY, BSDF_SPECULAR, BSDF_ALL = 1, 2, 4, 8, 16, 31


def _oracle_with(materials):
    """An oracle context whose material table holds the given BxDF lists (one dummy triangle carries the scene)."""
    sb = host.SceneBuilder()
    ids = [sb.material_lobes(l) for l in materials]
    sb.mesh([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 1, 2]], material=ids[0])
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    return o, ids


def _dirs(n, seed, hemisphere=None):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    if hemisphere == "+":
        v[:, 2] = np.abs(v[:, 2])
    # float32 directions, as the reference's Vectors are
    return v.astype(np.float32).astype(np.float64)


def _hemi_grid(n_th=400, n_ph=800, full_sphere=False):
    """Midpoint quadrature nodes and weights (sin theta dtheta dphi) over the +z hemisphere or the whole sphere."""
    tmax = np.pi if full_sphere else np.pi / 2
    th = (np.arange(n_th) + 0.5) / n_th * tmax
    ph = (np.arange(n_ph) + 0.5) / n_ph * (2 * np.pi)
    T, P = np.meshgrid(th, ph, indexing="ij")
    w = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], axis=-1).reshape(-1, 3)
    dw = (np.sin(T) * (tmax / n_th) * (2 * np.pi / n_ph)).reshape(-1)
    return w, dw


# ---- numpy restatements (from the reference's formulas) ----------------------------------------------------------------
def np_dielectric(cosi, etai, etat):  # fresnel_dielectric.dart:24-56
    cosi = np.clip(np.asarray(cosi, np.float64), -1.0, 1.0)
    entering = cosi > 0
    ei = np.where(entering, etai, etat)
    et = np.where(entering, etat, etai)
    sint = ei / et * np.sqrt(np.maximum(0.0, 1 - cosi * cosi))
    cost = np.sqrt(np.maximum(0.0, 1 - sint * sint))
    c = np.abs(cosi)
    rpar = (et * c - ei * cost) / (et * c + ei * cost)
    rper = (ei * c - et * cost) / (ei * c + et * cost)
    return np.where(sint >= 1.0, 1.0, 0.5 * (rpar * rpar + rper * rper))


def np_conductor(cosi, eta, k):  # fresnel_conductor.dart:28-45 (per channel)
    c = np.abs(np.asarray(cosi, np.float64))[..., None]
    eta, k = np.asarray(eta, np.float64), np.asarray(k, np.float64)
    tmp = (eta * eta + k * k) * c * c
    rpar2 = (tmp - 2 * eta * c + 1) / (tmp + 2 * eta * c + 1)
    tmpf = eta * eta + k * k
    rper2 = (tmpf - 2 * eta * c + c * c) / (tmpf + 2 * eta * c + c * c)
    return 0.5 * (rpar2 + rper2)


def np_blinn_microfacet(wo, wi, R, exponent, fresnel):  # microfacet.dart:27-56, blinn.dart:31-34
    co, ci = np.abs(wo[:, 2]), np.abs(wi[:, 2])
    wh = wi + wo
    nrm = np.linalg.norm(wh, axis=1, keepdims=True)
    ok = (co > 0) & (ci > 0) & (nrm[:, 0] > 0)
    wh = wh / np.where(nrm > 0, nrm, 1.0)
    cos_h = np.einsum("ij,ij->i", wi, wh)
    D = (exponent + 2.0) / (2 * np.pi) * np.abs(wh[:, 2]) ** exponent
    wo_h = np.abs(np.einsum("ij,ij->i", wo, wh))
    G = np.minimum(1.0, np.minimum(2 * np.abs(wh[:, 2]) * co / wo_h, 2 * np.abs(wh[:, 2]) * ci / wo_h))
    F = fresnel(cos_h)
    if F.ndim == 1:
        F = F[:, None]
    f = np.asarray(R, np.float64) * (D * G / (4 * ci * co))[:, None] * F
    return np.where(ok[:, None], f, 0.0)


def np_blinn_pdf(wo, wi, exponent):  # blinn.dart:61-71, microfacet.dart:67-72
    wh = wo + wi
    wh /= np.linalg.norm(wh, axis=1, keepdims=True)
    d = np.einsum("ij,ij->i", wo, wh)
    p = (exponent + 1.0) * np.abs(wh[:, 2]) ** exponent / (2 * np.pi * 4.0 * np.where(d > 0, d, 1.0))
    return np.where((d > 0) & (wo[:, 2] * wi[:, 2] > 0), p, 0.0)


def np_oren_nayar(wo, wi, R, sigma_deg):  # oren_nayar.dart:24-64
    s = math.radians(sigma_deg)
    A = 1 - s * s / (2 * (s * s + 0.33))
    B = 0.45 * s * s / (s * s + 0.09)
    sin_i, sin_o = np.sqrt(np.maximum(0, 1 - wi[:, 2] ** 2)), np.sqrt(np.maximum(0, 1 - wo[:, 2] ** 2))
    both = (sin_i > 1e-4) & (sin_o > 1e-4)
    si, so = np.where(both, sin_i, 1.0), np.where(both, sin_o, 1.0)
    dcos = np.clip(wi[:, 0] / si, -1, 1) * np.clip(wo[:, 0] / so, -1, 1) + np.clip(wi[:, 1] / si, -1, 1) * np.clip(wo[:, 1] / so, -1, 1)
    maxcos = np.where(both, np.maximum(0.0, dcos), 0.0)
    ci, co = np.abs(wi[:, 2]), np.abs(wo[:, 2])
    sinalpha = np.where(ci > co, sin_o, sin_i)
    tanbeta = np.where(ci > co, sin_i / ci, sin_o / co)
    return np.asarray(R, np.float64) * ((A + B * maxcos * sinalpha * tanbeta) / np.pi)[:, None]


# ---- f against the restatements, reciprocity ---------------------------------------------------------------------------
KD = np.array([0.7, 0.4, 0.2], np.float32)
KS = np.array([0.3, 0.5, 0.9], np.float32)
ETA = np.array([0.2, 0.9, 1.1], np.float32)
KK = np.array([3.9, 2.4, 2.2], np.float32)


@pytest.fixture(scope="module")
def zoo():
    mats = {
        "lambert": host.matte_lobes(KD),
        "oren": host.matte_lobes(KD, 35.0),
        "blinn_dielectric": host.plastic_lobes(0.0, KS, 0.15),
        "blinn_conductor": host.metal_lobes(ETA, KK, 0.08),
        "plastic": host.plastic_lobes(KD, KS, 0.15),
        "glass": host.glass_lobes(1.0, 1.0, 1.5),
        "mirror": host.mirror_lobes(0.8),
        "translucent": host.translucent_lobes(kd=0.8, ks=0.0, reflect=0.6, transmit=0.3),
        "mix": host.mix_lobes(host.matte_lobes(0.9), host.matte_lobes(0.2), amount=0.25),
        "shiny": host.shinymetal_lobes(ks=0.0, kr=(0.2, 0.7, 1.0)),
    }
    names = list(mats)
    o, ids = _oracle_with([mats[k] for k in names])
    return o, dict(zip(names, ids)), mats


def test_f_matches_the_numpy_restatement_of_every_bxdf(zoo):
    o, ids, mats = zoo
    wo, wi = _dirs(4000, 1, "+"), _dirs(4000, 2, "+")
    f, pdf = o.bsdf_eval(ids["lambert"], wo, wi)
    assert np.allclose(f, KD.astype(np.float64) / np.pi, rtol=2e-7)
    assert np.allclose(pdf, wi[:, 2] / np.pi, rtol=1e-12)  # bxdf.dart:84-88
    f, pdf = o.bsdf_eval(ids["oren"], wo, wi)
    assert np.allclose(f, np_oren_nayar(wo, wi, KD, 35.0), rtol=3e-7, atol=1e-9)
    assert np.allclose(pdf, wi[:, 2] / np.pi, rtol=1e-12)
    e = mats["blinn_dielectric"][0]["param"]
    assert e == pytest.approx(1.0 / 0.15)
    f, pdf = o.bsdf_eval(ids["blinn_dielectric"], wo, wi)
    expect = np_blinn_microfacet(wo, wi, KS, e, lambda c: np_dielectric(c, 1.5, 1.0))
    assert np.allclose(f, expect, rtol=2e-5, atol=1e-9)  # wh is a float32 Vector in the reference
    assert np.allclose(pdf, np_blinn_pdf(wo, wi, e), rtol=2e-5)
    e = mats["blinn_conductor"][0]["param"]
    f, pdf = o.bsdf_eval(ids["blinn_conductor"], wo, wi)
    expect = np_blinn_microfacet(wo, wi, 1.0, e, lambda c: np_conductor(c, ETA, KK))
    assert np.allclose(f, expect, rtol=2e-5, atol=1e-9)  # Spectrum arithmetic of the conductor term rounds to float32 per operation
    assert np.allclose(pdf, np_blinn_pdf(wo, wi, e), rtol=2e-5)
    # below the surface on the other side of the geometric normal nothing reflects (bsdf.dart:185-189)
    f, pdf = o.bsdf_eval(ids["plastic"], wo, wi * [1, 1, -1])
    assert (f == 0).all() and (pdf == 0).all()


@pytest.mark.parametrize("name", ["lambert", "oren", "blinn_dielectric", "blinn_conductor", "plastic", "mix"])
def test_helmholtz_reciprocity(zoo, name):
    o, ids, _ = zoo
    wo, wi = _dirs(3000, 3, "+"), _dirs(3000, 4, "+")
    f1, _ = o.bsdf_eval(ids[name], wo, wi)
    f2, _ = o.bsdf_eval(ids[name], wi, wo)
    assert np.allclose(f1, f2, rtol=2e-5, atol=1e-9)  # wh is a float32 Vector: cos(theta_h)^e carries e * 6e-8


# ---- energy, pdf normalisation, sampling consistency ----------------------------------------------------------------------
@pytest.mark.parametrize("name,cos_o", [(n, c) for n in ("lambert", "oren", "blinn_dielectric", "blinn_conductor", "plastic", "mix")
                                        for c in (1.0, 0.7, 0.25)])
def test_albedo_is_bounded_and_sample_f_is_consistent_with_f_and_pdf(zoo, name, cos_o):
    o, ids, _ = zoo
    wo = np.array([math.sqrt(1 - cos_o * cos_o), 0.0, cos_o], np.float32).astype(np.float64)
    wi, dw = _hemi_grid(500, 1000)
    f, pdf = o.bsdf_eval(ids[name], wo, wi)
    rho = (f.astype(np.float64) * (wi[:, 2] * dw)[:, None]).sum(axis=0)
    assert (rho <= 1.0 + 1e-3).all() and (rho > 0).all(), rho  # int f cos <= 1
    # the same integral by importance sampling through sample_f: a stratified grid of (u1, u2), component u3 cycling
    n = 256
    u1, u2 = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n, indexing="ij")
    u = np.stack([u1.ravel(), u2.ravel(), ((np.arange(n * n) * 0.6180339887) % 1.0)], axis=1)
    swi, sf, spdf, _ = o.bsdf_sample(ids[name], wo, u)
    ok = spdf > 0
    est = (sf[ok].astype(np.float64) * (np.abs(swi[ok, 2]) / spdf[ok])[:, None]).sum(axis=0) / (n * n)
    assert np.allclose(est, rho, rtol=2e-2), (est, rho)
    # the pdf sample_f reports is pdf() of the direction it returns, and f likewise.  (A Blinn sample that lands below the
    # horizon keeps the distribution's pdf with f = 0 in sample_f, microfacet.dart:58-65, while pdf() answers 0 there, :67-72:
    # harmless, the sample is worth nothing either way — compare where the sample carries something.)
    ok &= (sf != 0).any(axis=1)
    f2, pdf2 = o.bsdf_eval(ids[name], wo, swi[ok])
    assert np.allclose(pdf2, spdf[ok], rtol=2e-4)  # wi goes through float32 Vector stores in between (e * 6e-8 again)
    assert np.allclose(f2, sf[ok], rtol=2e-4, atol=1e-7)


def test_pdf_integrates_to_its_closed_form(zoo):
    o, ids, mats = zoo
    wi, dw = _hemi_grid(600, 600, full_sphere=True)
    up = np.array([0.0, 0.0, 1.0])
    for name in ("lambert", "oren", "mix"):  # the cosine density: 1 over wo's hemisphere, nothing on the other side
        _, pdf = o.bsdf_eval(ids[name], up, wi)
        assert (pdf * dw).sum() == pytest.approx(1.0, rel=1e-4)
        assert (pdf[wi[:, 2] < 0] == 0).all()
    # Blinn seen from the normal: wh ~ (e + 1) / (2 pi) cos^e, wi = reflect(wo, wh) stays above the horizon iff theta_h < 45 deg;
    # the mass of the rest is dropped by Microfacet.pdf (microfacet.dart:67-72): 1 - cos(45 deg)^(e + 1)
    e = mats["blinn_dielectric"][0]["param"]
    _, pdf = o.bsdf_eval(ids["blinn_dielectric"], up, wi)
    assert (pdf * dw).sum() == pytest.approx(1.0 - math.cos(math.pi / 4) ** (e + 1.0), rel=2e-3)
    # two matching BxDFs: the BSDF's pdf is their average (bsdf.dart:128-146)
    _, pdf = o.bsdf_eval(ids["plastic"], up, wi)
    assert (pdf * dw).sum() == pytest.approx(0.5 * (1.0 + 1.0 - math.cos(math.pi / 4) ** (e + 1.0)), rel=2e-3)


# ---- BSDF bookkeeping: component choice, wrappers ---------------------------------------------------------------------------
def test_component_choice_and_pdf_averaging(zoo):
    o, ids, mats = zoo
    wo = np.array([0.3, -0.2, math.sqrt(1 - 0.13)], np.float32).astype(np.float64)
    # plastic = [Lambertian, Microfacet]: component u < 0.5 samples the cosine lobe, u >= 0.5 the Blinn lobe (bsdf.dart:60-75)
    u = np.array([[0.3, 0.6, 0.49], [0.3, 0.6, 0.5]])
    wi, f, pdf, ty = o.bsdf_sample(ids["plastic"], wo, u)
    assert ty[0] == BSDF_REFLECTION | BSDF_DIFFUSE and ty[1] == BSDF_REFLECTION | BSDF_GLOSSY
    # CosineSampleHemisphere through ConcentricSampleDisk (montecarlo.dart:155-209) at (0.3, 0.6)
    sx, sy = 2 * np.float64(np.float32(0.3)) - 1, 2 * np.float64(np.float32(0.6)) - 1  # (-0.4, 0.2): third region, r = -sx, theta = 4 - sy / r
    r = -sx
    theta = (4.0 - sy / r) * (math.pi / 4)
    dx, dy = r * math.cos(theta), r * math.sin(theta)
    assert np.allclose(wi[0], [dx, dy, math.sqrt(max(0.0, 1 - dx * dx - dy * dy))], atol=2e-7)
    e = mats["plastic"][1]["param"]
    lam_pdf = abs(wi[0, 2]) / math.pi
    assert pdf[0] == pytest.approx(0.5 * (lam_pdf + np_blinn_pdf(wo[None], wi[:1].copy(), e)[0]), rel=1e-6)
    expect_f = KD / np.pi + np_blinn_microfacet(wo[None], wi[:1], KS, e, lambda c: np_dielectric(c, 1.5, 1.0))[0]
    assert np.allclose(f[0], expect_f, rtol=1e-6)
    # asking for the diffuse lobes only: one matching BxDF, no averaging, f without the glossy term
    wi_d, f_d, pdf_d, _ = o.bsdf_sample(ids["plastic"], wo, u[:1], flags=BSDF_REFLECTION | BSDF_DIFFUSE)
    assert np.allclose(wi_d, wi[:1]) and pdf_d[0] == pytest.approx(lam_pdf, rel=1e-6) and np.allclose(f_d[0], KD / np.pi, rtol=1e-6)


def test_brdf_to_btdf_and_scaled_bxdf(zoo):
    o, ids, mats = zoo
    wo, wi = _dirs(2000, 7, "+"), _dirs(2000, 8, "+")
    # translucent: Lambertian(r kd) on wo's side, BRDFToBTDF(Lambertian(t kd)) on the other (translucent_material.dart:62-70)
    f_r, pdf_r = o.bsdf_eval(ids["translucent"], wo, wi)
    f_t, pdf_t = o.bsdf_eval(ids["translucent"], wo, wi * [1, 1, -1])
    assert np.allclose(f_r, np.float32(0.6) * np.float32(0.8) / np.pi, rtol=1e-6)
    assert np.allclose(f_t, np.float32(0.3) * np.float32(0.8) / np.pi, rtol=1e-6)
    # both BxDFs match BSDF_ALL, so each side's pdf is half its cosine density
    assert np.allclose(pdf_r, 0.5 * wi[:, 2] / np.pi, rtol=1e-10) and np.allclose(pdf_t, pdf_r, rtol=1e-10)
    # mix: ScaledBxDF(Lambertian(.9), s) + ScaledBxDF(Lambertian(.2), 1 - s): f scales, pdf stays the cosine density
    f_m, pdf_m = o.bsdf_eval(ids["mix"], wo, wi)
    s = np.float32(0.25)
    expect = (np.float64(s * np.float32(0.9)) + np.float64(np.float32(1.0 - 0.25) * np.float32(0.2))) / np.pi
    assert np.allclose(f_m, expect, rtol=2e-6)
    assert np.allclose(pdf_m, wi[:, 2] / np.pi, rtol=1e-10)


# ---- specular lobes and the Fresnel terms -----------------------------------------------------------------------------------
def test_glass_splits_energy_and_follows_snell(zoo):
    o, ids, _ = zoo
    for cos_o in (1.0, 0.8, 0.3, -0.9, -0.5):  # negative: leaving the glass (total internal reflection at -0.5)
        wo = np.array([math.sqrt(1 - cos_o * cos_o), 0.0, cos_o], np.float32).astype(np.float64)
        # glass = [SpecularReflection, SpecularTransmission]: component < 0.5 reflects, >= 0.5 transmits; pdf 1 / 2 each
        wi, f, pdf, ty = o.bsdf_sample(ids["glass"], wo, np.array([[0.5, 0.5, 0.2], [0.5, 0.5, 0.7]]))
        F = float(np_dielectric(wo[2], 1.0, 1.5))
        assert ty[0] == BSDF_REFLECTION | BSDF_SPECULAR and pdf[0] == 0.5
        assert np.allclose(wi[0], [-wo[0], -wo[1], wo[2]])
        assert f[0, 0] * abs(wi[0, 2]) == pytest.approx(F, rel=1e-6)
        ei, et = (1.0, 1.5) if wo[2] > 0 else (1.5, 1.0)
        sint2 = (ei / et) ** 2 * (1 - wo[2] ** 2)
        if sint2 >= 1.0:
            assert F == 1.0 and pdf[1] == 0.0 and (f[1] == 0).all()  # total internal reflection (specular_transmission.dart:52-54)
            continue
        assert ty[1] == BSDF_TRANSMISSION | BSDF_SPECULAR and pdf[1] == 0.5
        assert f[1, 0] * abs(wi[1, 2]) == pytest.approx(1.0 - F, rel=1e-6)  # no eta^2 factor in the reference (:63-65)
        assert f[0, 0] * abs(wi[0, 2]) + f[1, 0] * abs(wi[1, 2]) == pytest.approx(1.0, rel=1e-6)
        # Snell: sin_t = (ei / et) sin_i, on the other side of the surface
        assert math.hypot(wi[1, 0], wi[1, 1]) == pytest.approx(ei / et * math.hypot(wo[0], wo[1]), rel=1e-6, abs=1e-9)
        assert wi[1, 2] * wo[2] < 0 and np.linalg.norm(wi[1]) == pytest.approx(1.0, rel=1e-6)
        # reciprocity of the Fresnel term: entering at theta_i reflects as much as leaving at theta_t
        assert float(np_dielectric(wi[1, 2], 1.0, 1.5)) == pytest.approx(F, rel=1e-5)
    # f() and pdf() of specular BxDFs are zero for every pair of directions
    f, pdf = o.bsdf_eval(ids["glass"], _dirs(100, 9), _dirs(100, 10))
    assert (f == 0).all() and (pdf == 0).all()


def test_conductor_fresnel_known_values(zoo):
    o, ids, _ = zoo
    # normal incidence: both polarisations give ((eta - 1)^2 + k^2) / ((eta + 1)^2 + k^2); grazing: 1
    R0 = ((ETA.astype(np.float64) - 1) ** 2 + KK.astype(np.float64) ** 2) / ((ETA.astype(np.float64) + 1) ** 2 + KK.astype(np.float64) ** 2)
    assert np.allclose(np_conductor(np.array([1.0]), ETA, KK)[0], R0, rtol=1e-12)
    assert np.allclose(np_conductor(np.array([0.0]), ETA, KK)[0], 1.0)
    # shinymetal's mirror lobe: FresnelApproxEta makes the normal-incidence reflectance equal to Kr (clamped at 0.999)
    wi, f, pdf, ty = o.bsdf_sample(ids["shiny"], np.array([0.0, 0.0, 1.0]), np.array([[0.1, 0.2, 0.3]]))
    assert ty[0] == BSDF_REFLECTION | BSDF_SPECULAR and pdf[0] == 1.0
    assert np.allclose(f[0], [0.2, 0.7, 0.999], rtol=2e-3)
    # the microfacet lobe of a metal in the mirror configuration: f = D(n) G F(cos) / (4 cos^2) with G = 1 below 45 degrees
    e = host.metal_lobes(ETA, KK, 0.08)[0]["param"]
    c = 0.9
    wo = np.array([math.sqrt(1 - c * c), 0.0, c])
    f, _ = o.bsdf_eval(ids["blinn_conductor"], wo, wo * [-1, -1, 1])
    assert np.allclose(f[0], (e + 2) / (2 * np.pi) * np_conductor(np.array([c]), ETA, KK)[0] / (4 * c * c), rtol=2e-6)


# ---- closed-box furnace through the integrators: pathLi / EstimateDirect against the numpy albedo -------------------------------
def _furnace_first_bounce(lobes, integ, spp=1024, radius=5.0):
    """A closed emissive sphere (Le = 1) seen from its centre: every direction of the hemisphere above any surface point sees the
    emitter, so the radiance after ONE bounce of direct lighting is  Le + Le * rho_hd(wo),  rho_hd(wo) = int f(wo, wi) |cos| dwi.
    The integrators get there by light sampling of the sphere (sphere.dart:247-311) + BSDF sampling with the power heuristic
    (integrator.dart:119-185); the albedo on the right comes from the numpy restatements above."""
    sb = host.SceneBuilder()
    sb.sphere(host.translate(0, 0, 0), radius=radius, material=sb.material_lobes(lobes), area_light=(1.0, 1.0, 1.0), reverse=True)
    cam = host.PerspectiveCamera(host.look_at((0, 0, 0), (0, 0, 1), (0, 1, 0)), fov=1.0)  # from the centre: wo = the normal
    o = Oracle()
    host.upload_scene(o, sb.arrays())
    host.configure_render(o, cam, host.Film(2, 2), host.Sampler(kind=host.SAMPLER_LD, spp=spp), integ)
    o.render(0, 1, 8)
    return o.film_read()["rgb"].astype(np.float64).mean(axis=(0, 1))


def _np_albedo(f_of_wi, n_th=600, n_ph=1200):
    wi, dw = _hemi_grid(n_th, n_ph)
    wo = np.broadcast_to(np.array([0.0, 0.0, 1.0]), wi.shape).copy()
    return (f_of_wi(wo, wi) * (wi[:, 2] * dw)[:, None]).sum(axis=0)


@pytest.mark.parametrize("integ", [host.Integrator(kind=host.INTEGRATOR_PATH, maxdepth=0), host.Integrator(kind=host.INTEGRATOR_DIRECT, maxdepth=1)],
                         ids=["path_depth0", "directlighting"])
def test_furnace_first_bounce_equals_one_plus_the_numpy_albedo(integ):
    cases = {
        "lambert": (host.matte_lobes(KD), lambda wo, wi: np.broadcast_to(KD.astype(np.float64) / np.pi, wi.shape)),
        "oren_nayar": (host.matte_lobes(KD, 35.0), lambda wo, wi: np_oren_nayar(wo, wi, KD, 35.0)),
        "plastic": (host.plastic_lobes(KD, KS, 0.15),
                    lambda wo, wi: KD.astype(np.float64) / np.pi + np_blinn_microfacet(wo, wi, KS, 1.0 / 0.15, lambda c: np_dielectric(c, 1.5, 1.0))),
        "metal": (host.metal_lobes(ETA, KK, 0.08),
                  lambda wo, wi: np_blinn_microfacet(wo, wi, 1.0, host.metal_lobes(ETA, KK, 0.08)[0]["param"], lambda c: np_conductor(c, ETA, KK))),
    }
    for name, (lobes, f) in cases.items():
        got = _furnace_first_bounce(lobes, integ)
        expect = 1.0 + _np_albedo(f)
        assert np.allclose(got, expect, rtol=1.5e-2), (name, got, expect)
