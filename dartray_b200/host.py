"""Host-side mirror of the reference objects the render path reads at the C-ABI boundary.

What a Dart `GpuSamplerRenderer` would pull out of the constructed DartRay objects,
computed here the way the reference computes it so that tests read like the reference's scene setup:

  Transform / Matrix4x4      lib/core/transform.dart:27-349, lib/core/matrix4x4.dart:26 (float32 storage)
  PerspectiveCamera          lib/cameras/perspective_camera.dart:46-57,134-181 + lib/core/projective_camera.dart:34-53
  filters                    lib/filters/*.dart (evaluate) -> ImageFilm's 16x16 table (lib/film/image_film.dart:74-82)
  Primitive.fullyRefine      lib/core/primitive.dart:71-84 (LIFO order) and ShapeSet (lib/core/light/shape_set.dart:26-41)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# quadric kinds of drt_set_quadrics (include/drt.h); sphere = 0 and disk = 1 have their own entry points
QUADRIC_CYLINDER, QUADRIC_CONE, QUADRIC_PARABOLOID, QUADRIC_HYPERBOLOID = 2, 3, 4, 5

# ---- Matrix4x4 / Transform (float32 storage, float64 arithmetic) ------------------------------------


def _m(a):
    return np.asarray(a, dtype=np.float64).reshape(4, 4).astype(np.float32)


def mat_mul(a, b):  # Matrix4x4.Mul, matrix4x4.dart
    return (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)


def mat_inv(a):
    return np.linalg.inv(a.astype(np.float64)).astype(np.float32)


def translate(x, y, z):  # transform.dart:214-228
    return _m([[1, 0, 0, x], [0, 1, 0, y], [0, 0, 1, z], [0, 0, 0, 1]])


def scale(x, y, z):  # transform.dart:230-243
    return _m([[x, 0, 0, 0], [0, y, 0, 0], [0, 0, z, 0], [0, 0, 0, 1]])


def rotate(angle_deg, axis):  # transform.dart:280-304
    a = np.asarray(axis, dtype=np.float64)
    a = (a / np.linalg.norm(a)).astype(np.float32).astype(np.float64)
    s, c = math.sin(math.radians(angle_deg)), math.cos(math.radians(angle_deg))
    m = np.zeros((4, 4))
    m[0] = [a[0] * a[0] + (1 - a[0] * a[0]) * c, a[0] * a[1] * (1 - c) - a[2] * s, a[0] * a[2] * (1 - c) + a[1] * s, 0]
    m[1] = [a[0] * a[1] * (1 - c) + a[2] * s, a[1] * a[1] + (1 - a[1] * a[1]) * c, a[1] * a[2] * (1 - c) - a[0] * s, 0]
    m[2] = [a[0] * a[2] * (1 - c) - a[1] * s, a[1] * a[2] * (1 - c) + a[0] * s, a[2] * a[2] + (1 - a[2] * a[2]) * c, 0]
    m[3] = [0, 0, 0, 1]
    return _m(m)


def look_at(pos, look, up):
    """Returns cameraToWorld (transform.dart:306-331 builds it as `m`; the CTM is its inverse)."""
    pos, look, up = (np.asarray(v, dtype=np.float64) for v in (pos, look, up))
    f32 = lambda v: v.astype(np.float32).astype(np.float64)
    d = f32((look - pos) / np.linalg.norm(look - pos))
    upn = f32(up / np.linalg.norm(up))
    left = np.cross(upn, d)
    left = f32(f32(left) / np.linalg.norm(f32(left)))
    new_up = f32(np.cross(d, left))
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, pos
    return _m(m)


def perspective(fov, n=1.0e-2, f=1000.0):  # transform.dart:338-349
    persp = _m([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, f / (f - n), -f * n / (f - n)], [0, 0, 1, 0]])
    inv_tan = 1.0 / math.tan(math.radians(fov) / 2.0)
    return mat_mul(scale(inv_tan, inv_tan, 1.0), persp)


def transform_points(m, pts):
    """Transform.transformPoint on float32 points (transform.dart:110-129)."""
    p = np.asarray(pts, dtype=np.float32).astype(np.float64).reshape(-1, 3)
    m64 = m.astype(np.float64)
    out = (p @ m64[:3, :3].T + m64[:3, 3]).astype(np.float32)
    w = p @ m64[3, :3] + m64[3, 3]
    ne = w != 1.0
    if ne.any():
        out[ne] = (out[ne].astype(np.float64) / w[ne, None]).astype(np.float32)
    return out


# ---- camera / film / sampler / integrator descriptions -------------------------------------------------


@dataclass
class PerspectiveCamera:
    camera_to_world: np.ndarray
    fov: float = 60.0
    lens_radius: float = 0.0
    focal_distance: float = 1.0e30
    shutter_open: float = 0.0
    shutter_close: float = 1.0
    screen_window: tuple | None = None

    def raster_to_camera(self, xres: int, yres: int) -> np.ndarray:
        frame = xres / yres  # perspective_camera.dart:150-171
        sw = self.screen_window
        if sw is None:
            sw = (-frame, frame, -1.0, 1.0) if frame > 1.0 else (-1.0, 1.0, -1.0 / frame, 1.0 / frame)
        screen_to_raster = mat_mul(mat_mul(scale(float(xres), float(yres), 1.0),
                                           scale(1.0 / (sw[1] - sw[0]), 1.0 / (sw[2] - sw[3]), 1.0)),
                                   translate(-sw[0], -sw[3], 0.0))  # projective_camera.dart:40-49
        raster_to_screen = mat_inv(screen_to_raster)
        return mat_mul(mat_inv(perspective(self.fov)), raster_to_screen)  # :51-52


def orthographic(znear=0.0, zfar=1.0):  # transform.dart:333-336
    return mat_mul(scale(1.0, 1.0, 1.0 / (zfar - znear)), translate(0.0, 0.0, -znear))


@dataclass
class OrthographicCamera(PerspectiveCamera):
    """Camera "orthographic" (lib/cameras/orthographic_camera.dart:38-50): Transform.Orthographic(0, 1) projection."""
    kind = 1

    def raster_to_camera(self, xres: int, yres: int) -> np.ndarray:
        frame = xres / yres
        sw = self.screen_window
        if sw is None:
            sw = (-frame, frame, -1.0, 1.0) if frame > 1.0 else (-1.0, 1.0, -1.0 / frame, 1.0 / frame)
        screen_to_raster = mat_mul(mat_mul(scale(float(xres), float(yres), 1.0),
                                           scale(1.0 / (sw[1] - sw[0]), 1.0 / (sw[2] - sw[3]), 1.0)),
                                   translate(-sw[0], -sw[3], 0.0))
        return mat_mul(mat_inv(orthographic(0.0, 1.0)), mat_inv(screen_to_raster))


@dataclass
class EnvironmentCamera(PerspectiveCamera):
    """Camera "environment" (lib/cameras/environment_camera.dart:37-52): no projection matrix, no lens."""
    kind = 2

    def raster_to_camera(self, xres: int, yres: int) -> np.ndarray:
        return np.eye(4, dtype=np.float32)


def filter_table(name: str = "box", xwidth: float | None = None, ywidth: float | None = None, **kw) -> tuple:
    """(xwidth, ywidth, float32[256]) — ImageFilm's precomputed table, image_film.dart:74-82."""
    defaults = {"box": 0.5, "gaussian": 2.0, "mitchell": 2.0, "triangle": 2.0, "sinc": 4.0}
    xw = defaults[name] if xwidth is None else xwidth
    yw = defaults[name] if ywidth is None else ywidth
    if name == "box":
        ev = lambda x, y: 1.0  # box_filter.dart:37
    elif name == "gaussian":  # gaussian_filter.dart:36-47
        a = kw.get("alpha", 2.0)
        ex, ey = math.exp(-a * xw * xw), math.exp(-a * yw * yw)
        ev = lambda x, y: max(0.0, math.exp(-a * x * x) - ex) * max(0.0, math.exp(-a * y * y) - ey)
    elif name == "mitchell":  # mitchell_filter.dart:41-55
        b, c = kw.get("B", 1.0 / 3.0), kw.get("C", 1.0 / 3.0)

        def m1(x):
            x = abs(2.0 * x)
            if x > 1.0:
                return ((-b - 6 * c) * x * x * x + (6 * b + 30 * c) * x * x + (-12 * b - 48 * c) * x + (8 * b + 24 * c)) * (1.0 / 6.0)
            return ((12 - 9 * b - 6 * c) * x * x * x + (-18 + 12 * b + 6 * c) * x * x + (6 - 2 * b)) * (1.0 / 6.0)

        ev = lambda x, y: m1(x * (1.0 / xw)) * m1(y * (1.0 / yw))
    elif name == "triangle":  # triangle_filter.dart:37
        ev = lambda x, y: max(0.0, xw - abs(x)) * max(0.0, yw - abs(y))
    elif name == "sinc":  # lanczos_sinc_filter.dart:39-55
        tau = kw.get("tau", 3.0)

        def s1(x):
            x = abs(x)
            if x < 1e-5:
                return 1.0
            if x > 1.0:
                return 0.0
            x *= math.pi
            return (math.sin(x) / x) * (math.sin(x * tau) / (x * tau))

        ev = lambda x, y: s1(x * (1.0 / xw)) * s1(y * (1.0 / yw))
    else:
        raise ValueError(name)
    t = np.empty(256, dtype=np.float32)
    for y in range(16):
        fy = (y + 0.5) * yw / 16
        for x in range(16):
            fx = (x + 0.5) * xw / 16
            t[y * 16 + x] = ev(fx, fy)
    return xw, yw, t


@dataclass
class Film:
    xres: int
    yres: int
    filter: str = "box"
    xwidth: float | None = None
    ywidth: float | None = None
    crop: tuple = (0.0, 1.0, 0.0, 1.0)
    filter_params: dict = field(default_factory=dict)

    def table(self):
        return filter_table(self.filter, self.xwidth, self.ywidth, **self.filter_params)

    def extent(self):  # image_film.dart:67-70
        left = math.ceil(self.xres * self.crop[0])
        width = max(1, math.ceil(self.xres * self.crop[1]) - left)
        top = math.ceil(self.yres * self.crop[2])
        height = max(1, math.ceil(self.yres * self.crop[3]) - top)
        return left, top, width, height


SAMPLER_LD, SAMPLER_STRATIFIED, SAMPLER_RANDOM, SAMPLER_HALTON, SAMPLER_ADAPTIVE, SAMPLER_BEST_CANDIDATE = 0, 1, 2, 3, 4, 5
ADAPTIVE_SHAPE_ID, ADAPTIVE_CONTRAST = 0, 1  # adaptive_sampler.dart:37-38: the `method` (carried in Sampler.jitter)
INTEGRATOR_PATH, INTEGRATOR_AO, INTEGRATOR_DIRECT, INTEGRATOR_WHITTED = 0, 1, 2, 3
RNG_SERIAL, RNG_KEYED = 0, 1


@dataclass
class Sampler:
    kind: int = SAMPLER_LD
    spp: int = 4          # lowdiscrepancy / random: pixelsamples
    xs: int = 2           # stratified: xsamples, ysamples; adaptive: minsamples, maxsamples
    ys: int = 2
    jitter: bool = True   # adaptive: the method, ADAPTIVE_SHAPE_ID / ADAPTIVE_CONTRAST
    pixel_order: int = 1  # 0 linear, 1 tile (render_options.dart default 'tile')
    tile_size: int = 32
    seed: int = 0
    rng_mode: int = RNG_KEYED
    sample_table: object = None  # bestcandidate: the 4096 x 5 pattern (best_candidate_sampler.dart:163-4258), handed over by the caller


@dataclass
class Integrator:
    kind: int = INTEGRATOR_DIRECT
    maxdepth: int = 5
    strategy: int = 0          # directlighting: 0 all, 1 one
    ao_nsamples: int = 2048
    ao_mindist: float = 1.0e-4
    ao_maxdist: float = math.inf


# ---- scene assembly: meshes / spheres / lights in the reference's refine order ---------------------------


# ---- materials as ordered BxDF lists (drt_set_material_lobes) ----------------------------------------------------
# The Dart shim does the same flattening when it walks GeometricPrimitive.material: every Material.getBSDF below only
# evaluates constant textures, so the BxDF list of a material is a constant of the scene.  Spectrum arithmetic is
# float32 storage / float64 expressions like RGBColor (rgb_color.dart:142-169).
LOBE_LAMBERTIAN, LOBE_OREN_NAYAR, LOBE_MICROFACET_BLINN, LOBE_SPECULAR_REFLECTION, LOBE_SPECULAR_TRANSMISSION, LOBE_FRESNEL_BLEND = range(6)
LOBE_REGULAR_HALFANGLE, LOBE_IRREGULAR_ISOTROPIC = 6, 7  # MeasuredMaterial's BxDFs: param = table index (drt_set_measured)
MEASURED_REGULAR_HALFANGLE, MEASURED_IRREGULAR_ISOTROPIC = 0, 1
FRESNEL_NOOP, FRESNEL_DIELECTRIC, FRESNEL_CONDUCTOR = range(3)


def _spec(v) -> np.ndarray:
    return np.broadcast_to(np.asarray(v, np.float64), (3,)).astype(np.float32)


def _clamp(s) -> np.ndarray:  # Spectrum.clamp (spectrum.dart:263-269)
    return np.clip(_spec(s).astype(np.float64), 0.0, np.inf).astype(np.float32)


def _mul(a, b) -> np.ndarray:
    return (a.astype(np.float64) * np.asarray(b, np.float64)).astype(np.float32)


def _black(s) -> bool:
    return not bool(np.any(s != 0))


# ---- textures: the constant-valued subset, folded on the host -----------------------------------------------------
# Material.getBSDF evaluates its textures per hit; ConstantTexture (core/texture/constant_texture.dart:23-37) ignores the hit, and so
# do ScaleTexture (textures/scale_texture.dart:26-34) and MixTexture (textures/mix_texture.dart:26-31) over constant inputs, so a tree
# of these three folds to one value when the scene is flattened.  Float textures are Dart doubles; spectrum textures round to float32
# after every operator like RGBColor (rgb_color.dart:136-151).  Anything that reads the hit point (imagemap, checkerboard, ...) is not
# on the GPU path yet: GpuUnsupported, as the Dart shim raises it (dart/lib/gpu/gpu_sampler_renderer.dart).
class GpuUnsupported(ValueError):
    pass


class Texture:
    def evaluate(self):
        raise GpuUnsupported(f"{type(self).__name__} is not a constant texture")


def _is_num(v) -> bool:
    return isinstance(v, (int, float, np.floating, np.integer))


def _tex_value(v):
    """One texture value as the reference holds it: a double, or a float32 RGB spectrum."""
    return float(v) if _is_num(v) else np.asarray(v, np.float64).reshape(3).astype(np.float32)


def _tex_mul(a, b):  # RGBColor.operator* with a num or a spectrum; double * double for float textures
    if _is_num(a) and _is_num(b):
        return a * b
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64)).astype(np.float32)


class ConstantTexture(Texture):
    def __init__(self, value=1.0):
        self.value = _tex_value(value)

    def evaluate(self):
        return self.value


class ScaleTexture(Texture):
    def __init__(self, tex1=1.0, tex2=1.0):
        self.tex1, self.tex2 = as_texture(tex1), as_texture(tex2)

    def evaluate(self):
        t1, t2 = self.tex1.evaluate(), self.tex2.evaluate()
        return _tex_mul(t2, t1) if _is_num(t1) else _tex_mul(t1, t2)


class MixTexture(Texture):
    def __init__(self, tex1=0.0, tex2=1.0, amount=0.5):
        self.tex1, self.tex2, self.amount = as_texture(tex1), as_texture(tex2), as_texture(amount)

    def evaluate(self):
        t1, t2, amt = self.tex1.evaluate(), self.tex2.evaluate(), self.amount.evaluate()
        if not _is_num(amt):
            raise ValueError("MixTexture amount is a float texture (mix_texture.dart:33-45)")
        a, b = _tex_mul(t1, 1.0 - amt), _tex_mul(t2, amt)
        if _is_num(a) and _is_num(b):
            return a + b
        if _is_num(a) or _is_num(b):
            raise ValueError("MixTexture mixes two float or two spectrum textures")
        return (a.astype(np.float64) + b.astype(np.float64)).astype(np.float32)


class OpaqueTexture(Texture):
    """Stands for a texture plugin that reads the hit point (imagemap, checkerboard, bilerp, ...): evaluate() raises."""

    def __init__(self, plugin: str):
        self.plugin = plugin

    def evaluate(self):
        raise GpuUnsupported(f"texture '{self.plugin}' is not a constant texture")


# ---- textures that read the hit point (drt_set_textures, include/drt.h) ---------------------------------
# numpy twins of the C structs drt_texture / drt_material_program
TEX_DTYPE = np.dtype([("kind", "i4"), ("spectrum", "i4"), ("tex1", "i4"), ("tex2", "i4"), ("amount", "i4"), ("mapping", "i4"),
                      ("image_width", "i4"), ("image_height", "i4"), ("image_channels", "i4"), ("image_wrap", "i4"),
                      ("image_trilinear", "i4"), ("aa_method", "i4"), ("image_offset", "u8"), ("value", "f8", (3,)),
                      ("value2", "f8", (9,)), ("su", "f8"), ("sv", "f8"), ("du", "f8"), ("dv", "f8"), ("max_anisotropy", "f8"),
                      ("world_to_texture", "f4", (16,)), ("v1", "f4", (3,)), ("v2", "f4", (3,))], align=True)
PROG_DTYPE = np.dtype([("kind", "i4"), ("tex", "i4", (8,)), ("bump", "i4"), ("m1", "i4"), ("m2", "i4")], align=True)
assert TEX_DTYPE.itemsize == 280 and PROG_DTYPE.itemsize == 48
WRAP_REPEAT, WRAP_BLACK, WRAP_CLAMP = 0, 1, 2  # mipmap.dart:24-26


@dataclass
class UVMapping:  # lib/core/texture/uv_mapping_2d.dart
    su: float = 1.0
    sv: float = 1.0
    du: float = 0.0
    dv: float = 0.0
    kind = 0


@dataclass
class SphericalMapping:  # spherical_mapping_2d.dart (Transform.Inverse(tex2world))
    world_to_texture: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    kind = 1


@dataclass
class CylindricalMapping(SphericalMapping):  # cylindrical_mapping_2d.dart
    kind = 2


@dataclass
class PlanarMapping:  # planar_mapping_2d.dart
    v1: tuple = (1.0, 0.0, 0.0)
    v2: tuple = (0.0, 1.0, 0.0)
    ds: float = 0.0
    dt: float = 0.0
    kind = 3


class HitPointTexture(Texture):
    """A texture whose value depends on the DifferentialGeometry: it cannot fold (evaluate() raises GpuUnsupported, which sends
    the material to SceneBuilder.material_program instead of the *_lobes functions)."""


class ImageTexture(HitPointTexture):
    """image_texture.dart.  `texels`: level 0 of the MIPMap as its constructor left it — (H, W) for a float texture, (H, W, 3) for a
    spectrum texture, power-of-two resolution, scale / gamma already applied (image_texture.dart:52-58)."""

    def __init__(self, texels, mapping=None, trilinear=False, max_anisotropy=8.0, wrap=WRAP_REPEAT):
        self.texels = np.ascontiguousarray(texels, np.float32)
        h, w = self.texels.shape[:2]
        if (w & (w - 1)) or (h & (h - 1)):
            raise ValueError("level 0 must have power-of-two resolution (the reference resamples at load, mipmap.dart:72-139)")
        self.mapping, self.trilinear, self.max_anisotropy, self.wrap = mapping or UVMapping(), trilinear, max_anisotropy, wrap


class CheckerboardTexture(HitPointTexture):  # checkerboard_texture.dart, dimension 2
    def __init__(self, tex1=1.0, tex2=0.0, mapping=None, aa="closedform"):
        self.tex1, self.tex2, self.mapping = as_texture(tex1), as_texture(tex2), mapping or UVMapping()
        self.aa = {"none": 0, "closedform": 1}[aa]


class UVTexture(HitPointTexture):  # uv_texture.dart
    def __init__(self, mapping=None):
        self.mapping = mapping or UVMapping()


class BilerpTexture(HitPointTexture):  # bilerp_texture.dart
    def __init__(self, v00=0.0, v01=1.0, v10=0.0, v11=1.0, mapping=None):
        self.v, self.mapping = (v00, v01, v10, v11), mapping or UVMapping()


class _Noise3D(HitPointTexture):
    """Textures over IdentityMapping3D (identity_mapping_3d.dart): `mapping_transform` is the Transform the mapping holds — the
    plugins' Create functions hand it tex2world as it is."""
    kind = 7

    def __init__(self, octaves=8, roughness=0.5, mapping_transform=None):
        self.octaves, self.roughness = int(octaves), float(roughness)
        self.w2t = np.eye(4, dtype=np.float32) if mapping_transform is None else mapping_transform


class FBmTexture(_Noise3D):  # fbm_texture.dart
    kind = 7


class WrinkledTexture(_Noise3D):  # wrinkled_texture.dart
    kind = 8


class WindyTexture(_Noise3D):  # windy_texture.dart
    kind = 9

    def __init__(self, mapping_transform=None):
        super().__init__(0, 0.0, mapping_transform)


class MarbleTexture(_Noise3D):  # marble_texture.dart (spectrum only)
    kind = 10

    def __init__(self, octaves=8, roughness=0.5, scale=1.0, variation=0.2, mapping_transform=None):
        super().__init__(octaves, roughness, mapping_transform)
        self.scale, self.variation = float(scale), float(variation)


class DotsTexture(HitPointTexture):  # dots_texture.dart
    def __init__(self, inside=1.0, outside=0.0, mapping=None):
        self.inside, self.outside, self.mapping = as_texture(inside), as_texture(outside), mapping or UVMapping()


class Checkerboard3DTexture(HitPointTexture):  # checkerboard_3d_texture.dart
    def __init__(self, tex1=1.0, tex2=0.0, mapping_transform=None):
        self.tex1, self.tex2 = as_texture(tex1), as_texture(tex2)
        self.w2t = np.eye(4, dtype=np.float32) if mapping_transform is None else mapping_transform


def is_constant_texture(v) -> bool:
    if not isinstance(v, Texture):
        return True
    try:
        v.evaluate()
        return True
    except GpuUnsupported:
        return False


class TextureTable:
    """Flattens texture trees into the node array of drt_set_textures (children before parents, one node per (object, type))."""

    def __init__(self):
        self.nodes, self.texels, self._ids, self._ntex = [], [], {}, 0

    def _mapping(self, n, m):
        n["mapping"] = m.kind
        if m.kind == 0:
            n["su"], n["sv"], n["du"], n["dv"] = m.su, m.sv, m.du, m.dv
        elif m.kind in (1, 2):
            n["world_to_texture"] = _m(m.world_to_texture).reshape(16)
        else:
            n["v1"], n["v2"], n["du"], n["dv"] = m.v1, m.v2, m.ds, m.dt

    def add(self, v, spectrum: bool) -> int:
        key = (id(v), spectrum) if isinstance(v, Texture) else None
        if key in self._ids:
            return self._ids[key]
        n = np.zeros((), TEX_DTYPE)
        n["spectrum"] = int(spectrum)
        n["tex1"] = n["tex2"] = n["amount"] = -1
        n["su"] = n["sv"] = 1.0
        n["max_anisotropy"] = 8.0
        n["world_to_texture"] = np.eye(4, dtype=np.float32).reshape(16)

        def const(val):
            n["kind"] = 0
            n["value"] = np.broadcast_to(np.asarray(val, np.float64), (3,)) if spectrum else (float(np.asarray(val).reshape(-1)[0]), 0, 0)

        if not isinstance(v, Texture):
            const(v)
        elif is_constant_texture(v):
            const(v.evaluate())  # a constant tree folds here exactly as fold_texture folds it
        elif isinstance(v, ScaleTexture):
            n["kind"], n["tex1"], n["tex2"] = 1, self.add(v.tex1, spectrum), self.add(v.tex2, spectrum)
        elif isinstance(v, MixTexture):
            n["kind"], n["tex1"], n["tex2"], n["amount"] = 2, self.add(v.tex1, spectrum), self.add(v.tex2, spectrum), self.add(v.amount, False)
        elif isinstance(v, ImageTexture):
            ch = 3 if v.texels.ndim == 3 else 1
            if ch != (3 if spectrum else 1):
                raise ValueError("a spectrum parameter needs an (H, W, 3) image, a float parameter an (H, W) image (image_texture.dart:38-42)")
            n["kind"], n["image_height"], n["image_width"], n["image_channels"] = 3, v.texels.shape[0], v.texels.shape[1], ch
            n["image_wrap"], n["image_trilinear"], n["max_anisotropy"] = v.wrap, int(v.trilinear), v.max_anisotropy
            n["image_offset"] = self._ntex
            self.texels.append(v.texels.reshape(-1))
            self._ntex += v.texels.size
            self._mapping(n, v.mapping)
        elif isinstance(v, CheckerboardTexture):
            n["kind"], n["tex1"], n["tex2"], n["aa_method"] = 4, self.add(v.tex1, spectrum), self.add(v.tex2, spectrum), v.aa
            self._mapping(n, v.mapping)
        elif isinstance(v, UVTexture):
            if not spectrum:
                raise GpuUnsupported("'uv' has no float form (uv_texture.dart:39-41)")
            n["kind"] = 5
            self._mapping(n, v.mapping)
        elif isinstance(v, _Noise3D):
            if v.kind == 10 and not spectrum:
                raise GpuUnsupported("'marble' has no float form (marble_texture.dart:68-70)")
            n["kind"], n["aa_method"], n["mapping"] = v.kind, v.octaves, 4
            n["value"] = (v.roughness, getattr(v, "scale", 0.0), getattr(v, "variation", 0.0))
            n["world_to_texture"] = _m(v.w2t).reshape(16)
        elif isinstance(v, DotsTexture):
            n["kind"], n["tex1"], n["tex2"] = 11, self.add(v.outside, spectrum), self.add(v.inside, spectrum)
            self._mapping(n, v.mapping)
        elif isinstance(v, Checkerboard3DTexture):
            n["kind"], n["tex1"], n["tex2"], n["mapping"] = 12, self.add(v.tex1, spectrum), self.add(v.tex2, spectrum), 4
            n["world_to_texture"] = _m(v.w2t).reshape(16)
        elif isinstance(v, BilerpTexture):
            n["kind"] = 6
            vals = [np.broadcast_to(np.asarray(x, np.float64), (3,)) for x in v.v]
            n["value"], n["value2"] = vals[0], np.concatenate(vals[1:])
            self._mapping(n, v.mapping)
        else:
            raise GpuUnsupported(f"texture {type(v).__name__} is not on the GPU path")
        self.nodes.append(n)
        if key:
            self._ids[key] = len(self.nodes) - 1
        return len(self.nodes) - 1

    def arrays(self):
        nodes = np.asarray(self.nodes, TEX_DTYPE) if self.nodes else np.zeros(0, TEX_DTYPE)
        return nodes, (np.concatenate(self.texels) if self.texels else np.zeros(0, np.float32))


# parameter slots of drt_material_program per material plugin: (name, is spectrum, default) in tex[] order (include/drt.h)
PROGRAM_PARAMS = {
    "matte": (0, (("kd", True, 0.5), ("sigma", False, 0.0))),
    "mirror": (1, (("kr", True, 0.9),)),
    "glass": (2, (("kr", True, 1.0), ("kt", True, 1.0), ("index", False, 1.5))),
    "plastic": (3, (("kd", True, 0.25), ("ks", True, 0.25), ("roughness", False, 0.1))),
    "metal": (4, (("eta", True, None), ("k", True, None), ("roughness", False, 0.01))),
    "shinymetal": (5, (("ks", True, 1.0), ("kr", True, 1.0), ("roughness", False, 0.1))),
    "substrate": (6, (("kd", True, 0.5), ("ks", True, 0.5), ("uroughness", False, 0.1), ("vroughness", False, 0.1))),
    "translucent": (7, (("kd", True, 0.25), ("ks", True, 0.25), ("reflect", True, 0.5), ("transmit", True, 0.5), ("roughness", False, 0.1))),
    "uber": (8, (("kd", True, 0.25), ("ks", True, 0.25), ("kr", True, 0.0), ("kt", True, 0.0), ("roughness", False, 0.1),
                 ("opacity", True, 1.0), ("index", False, 1.5))),
    "mix": (9, (("amount", True, 0.5),)),
    "subsurface": (10, (("kr", True, 1.0), ("index", False, 1.3))),
    "kdsubsurface": (10, (("kr", True, 1.0), ("index", False, 1.3))),
    "measured": (11, ()),  # m1 = SceneBuilder.measured_table(...)
}


def as_texture(v) -> Texture:
    return v if isinstance(v, Texture) else ConstantTexture(v)


def fold_texture(v):
    """The value a material parameter takes at every hit: numbers and RGB triples pass through, texture trees fold."""
    return v.evaluate() if isinstance(v, Texture) else v


def _folds_textures(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kw):
        return fn(*[fold_texture(a) for a in args], **{k: fold_texture(v) for k, v in kw.items()})
    return wrapper


WRAP_BTDF, WRAP_SCALED = 1, 2  # drt_set_lobe_wrappers: BRDFToBTDF(bxdf), ScaledBxDF(.., scale)


def _lobe(kind, rgb, fresnel=FRESNEL_NOOP, eta=(0, 0, 0), k=(0, 0, 0), param=0.0, ei=1.0, et=1.0, wrap=0, scale=(1, 1, 1)) -> dict:
    return dict(kind=int(kind), rgb=_spec(rgb), fresnel=int(fresnel), eta=_spec(eta), k=_spec(k), param=float(param),
                ei=float(ei), et=float(et), wrap=int(wrap), scale=_spec(scale))


def _blinn_exponent(roughness: float) -> float:  # 1 / roughness, then blinn.dart:24-28
    with np.errstate(divide="ignore"):
        e = float(np.float64(1.0) / np.float64(roughness))
    return 10000.0 if (e > 10000.0 or math.isnan(e)) else e


@_folds_textures
def matte_lobes(kd=0.5, sigma=0.0) -> list:  # matte_material.dart:41-65
    r, sig = _clamp(kd), min(max(float(sigma), 0.0), 90.0)
    if _black(r):
        return []
    return [_lobe(LOBE_LAMBERTIAN, r)] if sig == 0.0 else [_lobe(LOBE_OREN_NAYAR, r, param=sig)]


@_folds_textures
def mirror_lobes(kr=0.9) -> list:  # mirror_material.dart:26-43
    r = _clamp(kr)
    return [] if _black(r) else [_lobe(LOBE_SPECULAR_REFLECTION, r, FRESNEL_NOOP)]


@_folds_textures
def glass_lobes(kr=1.0, kt=1.0, index=1.5) -> list:  # glass_material.dart:26-52
    out, r, t = [], _clamp(kr), _clamp(kt)
    if not _black(r):
        out.append(_lobe(LOBE_SPECULAR_REFLECTION, r, FRESNEL_DIELECTRIC, ei=1.0, et=index))
    if not _black(t):
        out.append(_lobe(LOBE_SPECULAR_TRANSMISSION, t, FRESNEL_DIELECTRIC, ei=1.0, et=index))
    return out


@_folds_textures
def plastic_lobes(kd=0.25, ks=0.25, roughness=0.1) -> list:  # plastic_material.dart:26-53
    out, d, sp = [], _clamp(kd), _clamp(ks)
    if not _black(d):
        out.append(_lobe(LOBE_LAMBERTIAN, d))
    if not _black(sp):
        out.append(_lobe(LOBE_MICROFACET_BLINN, sp, FRESNEL_DIELECTRIC, param=_blinn_exponent(roughness), ei=1.5, et=1.0))
    return out


def _f32(a):
    return np.asarray(a, np.float64).astype(np.float32)


@_folds_textures
def shinymetal_lobes(ks=1.0, kr=1.0, roughness=0.1) -> list:  # shiny_metal_material.dart:42-76
    out = []

    def approx_eta(fr):  # FresnelApproxEta (:66-70): Spectrum operations round to float32 one by one
        refl = np.clip(_clamp(fr).astype(np.float64), 0.0, 0.999).astype(np.float32)
        sq = _f32(np.sqrt(refl.astype(np.float64)))
        return _f32(_f32(1.0 + sq.astype(np.float64)).astype(np.float64) / _f32(1.0 - sq.astype(np.float64)).astype(np.float64))
    sp, r = _clamp(ks), _clamp(kr)
    if not _black(sp):
        out.append(_lobe(LOBE_MICROFACET_BLINN, 1.0, FRESNEL_CONDUCTOR, eta=approx_eta(sp), k=(0, 0, 0), param=_blinn_exponent(roughness)))
    if not _black(r):
        out.append(_lobe(LOBE_SPECULAR_REFLECTION, 1.0, FRESNEL_CONDUCTOR, eta=approx_eta(r), k=(0, 0, 0)))
    return out


@_folds_textures
def translucent_lobes(kd=0.25, ks=0.25, reflect=0.5, transmit=0.5, roughness=0.1) -> list:  # translucent_material.dart:47-90
    out, r, t = [], _clamp(reflect), _clamp(transmit)
    if _black(r) and _black(t):
        return out
    d = _clamp(kd)
    if not _black(d):
        if not _black(r):
            out.append(_lobe(LOBE_LAMBERTIAN, _mul(r, d)))
        if not _black(t):
            out.append(_lobe(LOBE_LAMBERTIAN, _mul(t, d), wrap=WRAP_BTDF))
    sp = _clamp(ks)
    if not _black(sp):
        e = _blinn_exponent(roughness)
        if not _black(r):
            out.append(_lobe(LOBE_MICROFACET_BLINN, _mul(r, sp), FRESNEL_DIELECTRIC, param=e, ei=1.5, et=1.0))
        if not _black(t):
            out.append(_lobe(LOBE_MICROFACET_BLINN, _mul(t, sp), FRESNEL_DIELECTRIC, param=e, ei=1.5, et=1.0, wrap=WRAP_BTDF))
    return out


def mix_lobes(lobes1, lobes2, amount=0.5) -> list:  # mix_material.dart:36-50
    s1 = _clamp(fold_texture(amount))
    s2 = np.clip(_f32(1.0 - s1.astype(np.float64)).astype(np.float64), 0.0, np.inf).astype(np.float32)
    out = []
    for ll, sc in ((lobes1, s1), (lobes2, s2)):
        for l in ll:
            if l["wrap"] & WRAP_SCALED:
                raise ValueError("a mix of mixes (ScaledBxDF of a ScaledBxDF) is not representable in drt_set_lobe_wrappers")
            m = dict(l)
            m["wrap"] = l["wrap"] | WRAP_SCALED
            m["scale"] = sc
            out.append(m)
    if len(out) > 8:
        raise ValueError("a BSDF holds at most 8 BxDFs (bsdf.dart:253)")
    return out


@_folds_textures
def substrate_lobes(kd=0.5, ks=0.5, uroughness=0.1, vroughness=0.1) -> list:  # substrate_material.dart:46-68
    d, sp = _clamp(kd), _clamp(ks)
    if _black(d) and _black(sp):
        return []
    # FresnelBlend(d, s, Anisotropic(1 / u, 1 / v)): Rs travels in the eta slot, the two exponents (clamped as anisotropic.dart:30-37
    # clamps them, the same rule as Blinn's) in param / ei
    return [_lobe(LOBE_FRESNEL_BLEND, d, eta=sp, param=_blinn_exponent(uroughness), ei=_blinn_exponent(vroughness))]


@_folds_textures
def metal_lobes(eta, k, roughness=0.01) -> list:  # metal_material.dart:26-46 (eta / k given as RGB)
    return [_lobe(LOBE_MICROFACET_BLINN, 1.0, FRESNEL_CONDUCTOR, eta=eta, k=k, param=_blinn_exponent(roughness))]


@_folds_textures
def subsurface_lobes(kr=1.0, index=1.3) -> list:  # subsurface_material.dart:52-69, kd_subsurface_material.dart:48-67
    """The BSDF of both subsurface materials is one SpecularReflection(Kr, FresnelDielectric(1, index)); their BSSRDF is read by the
    dipole integrator only, which is not on the path."""
    r = _clamp(kr)
    return [] if _black(r) else [_lobe(LOBE_SPECULAR_REFLECTION, r, FRESNEL_DIELECTRIC, ei=1.0, et=index)]


def measured_lobes(table: int, table_kind: int) -> list:  # measured_material.dart:219-238
    """MeasuredMaterial.getBSDF without a bump map: one RegularHalfangleBRDF or IrregularIsotropicBRDF over table `table`."""
    return [_lobe(LOBE_REGULAR_HALFANGLE if table_kind == MEASURED_REGULAR_HALFANGLE else LOBE_IRREGULAR_ISOTROPIC, 1.0, param=table)]


def merl_table(data: bytes) -> np.ndarray:  # measured_material.dart:160-202
    """regularHalfangleData of a .merl file: three little-endian int32 dimensions, then one plane of float64 per colour channel; every
    value is rounded to float32 (the Float32List `tmp`), scaled by (1, 1.15, 1.66) / 1500 in binary64 and clamped below at 0."""
    dims = np.frombuffer(data, "<i4", 3)
    n = int(dims[0]) * int(dims[1]) * int(dims[2])
    if n != 90 * 90 * 180:
        raise ValueError("Dimensions don't match")
    raw = np.frombuffer(data, "<f8", 3 * n, 12).reshape(3, n).astype(np.float32).astype(np.float64)
    scales = np.array([1.0 / 1500.0, 1.15 / 1500.0, 1.66 / 1500.0])
    out = np.maximum(0.0, raw * scales[:, None]).astype(np.float32)
    return np.ascontiguousarray(out.T).reshape(90, 90, 180, 3)


def brdf_remap(wo, wi) -> np.ndarray:  # brdf_remap.dart:23-47 (float32 Vector components in, float32 Point out)
    wo, wi = np.asarray(wo, np.float32).astype(np.float64), np.asarray(wi, np.float32).astype(np.float64)
    cosi, coso = wi[..., 2], wo[..., 2]
    sini, sino = np.sqrt(np.maximum(0.0, 1.0 - cosi * cosi)), np.sqrt(np.maximum(0.0, 1.0 - coso * coso))
    phi = lambda v: np.where(np.arctan2(v[..., 1], v[..., 0]) < 0.0, np.arctan2(v[..., 1], v[..., 0]) + 2.0 * np.pi, np.arctan2(v[..., 1], v[..., 0]))
    dphi = phi(wi) - phi(wo)
    dphi = np.where(dphi < 0.0, dphi + 2.0 * np.pi, dphi)
    dphi = np.where(dphi > 2.0 * np.pi, dphi - 2.0 * np.pi, dphi)
    dphi = np.where(dphi > np.pi, 2.0 * np.pi - dphi, dphi)
    return np.stack([sini * sino, dphi / np.pi, cosi * coso], -1).astype(np.float32)


def brdf_samples(thetai, phii, thetao, phio, rgb) -> np.ndarray:  # measured_material.dart:117-131
    """The IrregIsotropicBRDFSamples of a .brdf file's measurements: the four angles of every measurement and its value ALREADY
    converted to RGB (Spectrum.fromSampled stays with the caller) -> n x 6 float32 (BRDFRemap point, RGB) in file order."""
    def sph(theta, phi):  # Vector.SphericalDirection(sin(theta), cos(theta), phi)
        st, ct, ph = np.sin(np.asarray(theta, np.float64)), np.cos(np.asarray(theta, np.float64)), np.asarray(phi, np.float64)
        return np.stack([st * np.cos(ph), st * np.sin(ph), ct], -1).astype(np.float32)
    p = brdf_remap(sph(thetao, phio), sph(thetai, phii))
    return np.concatenate([p, np.asarray(rgb, np.float32).reshape(-1, 3)], 1)


@_folds_textures
def uber_lobes(kd=0.25, ks=0.25, kr=0.0, kt=0.0, roughness=0.1, index=1.5, opacity=1.0) -> list:  # uber_material.dart:27-75
    out, op = [], _clamp(opacity)
    if not bool(np.all(op == 1.0)):
        t = ((-op.astype(np.float64)).astype(np.float32).astype(np.float64) + 1.0).astype(np.float32)
        out.append(_lobe(LOBE_SPECULAR_TRANSMISSION, t, FRESNEL_DIELECTRIC, ei=1.0, et=1.0))
    d = _mul(op, _clamp(kd))
    if not _black(d):
        out.append(_lobe(LOBE_LAMBERTIAN, d))
    sp = _mul(op, _clamp(ks))
    if not _black(sp):
        out.append(_lobe(LOBE_MICROFACET_BLINN, sp, FRESNEL_DIELECTRIC, param=_blinn_exponent(roughness), ei=index, et=1.0))
    r = _mul(op, _clamp(kr))
    if not _black(r):
        out.append(_lobe(LOBE_SPECULAR_REFLECTION, r, FRESNEL_DIELECTRIC, ei=index, et=1.0))
    t = _mul(op, _clamp(kt))
    if not _black(t):
        out.append(_lobe(LOBE_SPECULAR_TRANSMISSION, t, FRESNEL_DIELECTRIC, ei=index, et=1.0))
    return out


class SceneBuilder:
    """Collects shapes in scene-file order and produces the flat arrays of the C ABI."""

    def __init__(self):
        self.P, self.idx = [], []
        self.tri_mat, self.tri_light, self.tri_rev = [], [], []
        self.sph = []  # (o2w, w2o, params, mat, light, rev)
        self.dsk = []  # (o2w, w2o, (height, radius, innerradius, phimax), mat, light, rev)
        self.mesh_info, self.vertN, self.vertS, self.vertUV, self.tri_mesh = [], [], [], [], []  # per-vertex shading attributes
        self.quad = []  # (kind 2..5, o2w, w2o, 8 params, mat, light, rev): cylinder / cone / paraboloid / hyperboloid
        self.materials = []  # (kind, kd, sigma)
        self.lights = []  # dict(kind, L, pos, nsamples, shapes=[("tri"|"sph", local ids...)])
        self._order = []  # ("mesh", first_tri, ntris) | ("sphere", sphere_index)
        # TransformedPrimitives (transformed_primitive.dart): objects = the aggregates they wrap (each an order list like _order plus the
        # nested accelerator's parameters), instances = (object, world-to-primitive m / mInv at the start and end time, the two times)
        self._top_order = self._order
        self.objects = []
        self.instances = []
        self._nverts = 0

    def material(self, kd, sigma=0.0) -> int:
        kd, sigma = fold_texture(kd), fold_texture(sigma)  # constant Kd / sigma textures (matte_material.dart:41-65)
        self.materials.append((0, tuple(float(v) for v in np.broadcast_to(np.asarray(kd, np.float64), (3,))), float(sigma)))
        return len(self.materials) - 1

    def material_lobes(self, lobes: list) -> int:
        """A material given as its ordered BxDF list (mirror_lobes, glass_lobes, plastic_lobes, metal_lobes, uber_lobes ...)."""
        self.materials.append(("lobes", list(lobes)))
        return len(self.materials) - 1

    def measured_table(self, kind: int, data) -> int:
        """Data of one MeasuredMaterial file (drt_set_measured): kind MEASURED_REGULAR_HALFANGLE with an (nThetaH, nThetaD, nPhiD, 3)
        array (merl_table) or MEASURED_IRREGULAR_ISOTROPIC with an (n, 6) array (brdf_samples).  Returns the table index that
        measured_lobes / material_program("measured", m1=...) take."""
        if not hasattr(self, "measured"):
            self.measured = []
        a = np.ascontiguousarray(data, np.float32)
        if (kind == MEASURED_REGULAR_HALFANGLE and (a.ndim != 4 or a.shape[3] != 3)) or (kind == MEASURED_IRREGULAR_ISOTROPIC and (a.ndim != 2 or a.shape[1] != 6)):
            raise ValueError("measured table shape")
        self.measured.append((int(kind), a))
        return len(self.measured) - 1

    # -- TransformedPrimitive: object instancing and animated shapes (dartray.dart:404-452,480-546) -------------------------------
    def begin_object(self) -> None:
        """ObjectBegin: the shapes added until end_object() go into the object, not into the scene.  Their transforms stay their own
        (a shape inside an object is built with the CTM of its definition, dartray.dart:378-403)."""
        if self._order is not self._top_order:
            raise ValueError("ObjectBegin inside an object")
        self._order = []

    def end_object(self, split: int = 2, max_node_prims: int = 4) -> int:
        """ObjectEnd; returns the object index for instance().  split / max_node_prims: the accelerator objectInstance() builds
        over the object's primitives when it has more than one (the scene's own accelerator parameters, dartray.dart:520-536)."""
        if self._order is self._top_order:
            raise ValueError("ObjectEnd without ObjectBegin")
        if not self._order:
            raise ValueError("an empty object cannot be instanced (dartray.dart:514-516 returns)")
        self.objects.append((self._order, int(split), int(max_node_prims)))
        self._order = self._top_order
        return len(self.objects) - 1

    def instance(self, obj: int, instance_to_world, instance_to_world_end=None, start_time: float = 0.0, end_time: float = 1.0) -> int:
        """ObjectInstance under the CTM `instance_to_world` (and, inside an ActiveTransform / TransformTimes block, the end-time CTM):
        TransformedPrimitive(object, AnimatedTransform(Inverse(ctm0), t0, Inverse(ctm1), t1)) (dartray.dart:537-546)."""
        m0 = np.asarray(instance_to_world, np.float32).reshape(4, 4)
        m1 = m0 if instance_to_world_end is None else np.asarray(instance_to_world_end, np.float32).reshape(4, 4)
        # Transform.Inverse swaps m and mInv (transform.dart:58-60): world-to-primitive = (mInv, m) of the CTM
        self.instances.append((int(obj), mat_inv(m0), m0, mat_inv(m1), m1, float(start_time), float(end_time)))
        self._top_order.append(("instance", len(self.instances) - 1))
        return len(self.instances) - 1

    def animated(self, add_shape, o2w_start, o2w_end, start_time: float = 0.0, end_time: float = 1.0, max_node_prims: int = 1) -> int:
        """DartRay.shape for an animated CTM (dartray.dart:404-452): the shape is built under the IDENTITY transform — call
        `add_shape(sb)` to add it, without an o2w — wrapped in a BVHAccel with the constructor's defaults (maxPrims 1, sah) when it
        refines into more than one primitive, and instanced with AnimatedTransform(Inverse(ctm0), Inverse(ctm1))."""
        self.begin_object()
        add_shape(self)
        obj = self.end_object(split=2, max_node_prims=max_node_prims)
        return self.instance(obj, o2w_start, o2w_end, start_time, end_time)

    def material_program(self, plugin: str, bumpmap=None, m1=None, m2=None, **params) -> int:
        """A material whose parameters are textures that read the hit point, or that carries a bump map (`plugin` and the parameter
        names are the reference's, lower case: material_program("uber", kd=ImageTexture(...), ks=0.05, bumpmap=...)); "mix" takes
        the two material indices m1 / m2 and `amount`."""
        kind, slots = PROGRAM_PARAMS[plugin]
        unknown = set(params) - {n for n, _, _ in slots}
        if unknown:
            raise ValueError(f"{plugin} has no parameter {sorted(unknown)}")
        vals = []
        for name, spectrum, default in slots:
            v = params.get(name, default)
            if v is None:
                raise ValueError(f"{plugin} needs {name}")
            vals.append((v, spectrum))
        if kind == 9 and (m1 is None or m2 is None):
            raise ValueError("mix needs m1 and m2")
        if kind == 11 and m1 is None:
            raise ValueError("measured needs m1 = the index measured_table() returned")
        self.materials.append(("program", kind, vals, bumpmap, m1, m2))
        return len(self.materials) - 1

    # -- participating media (lib/volume_regions/*.dart, Volume "homogeneous" | "exponential" | "volumegrid") ------------------
    def volume(self, kind="homogeneous", sigma_a=0.0, sigma_s=0.0, g=0.0, le=0.0, p0=(0, 0, 0), p1=(1, 1, 1), volume_to_world=None,
               a=1.0, b=1.0, updir=(0, 1, 0), density=None) -> int:
        """One VolumeRegion with the reference's parameter names and defaults (homogenous_volume_region.dart:75-87,
        exponential_density_region.dart:52-66, volume_grid.dart:75-102).  `density`: (nz, ny, nx) array for "volumegrid"."""
        kinds = {"homogeneous": 0, "exponential": 1, "volumegrid": 2}
        rgb = lambda v: tuple(float(x) for x in np.broadcast_to(np.asarray(fold_texture(v), np.float64), (3,)))
        v2w = np.eye(4, dtype=np.float32) if volume_to_world is None else np.asarray(volume_to_world, np.float32).reshape(4, 4)
        d = None
        if kinds[kind] == 2:
            d = np.ascontiguousarray(density, np.float64)
            if d.ndim != 3:
                raise ValueError("volumegrid: density must be a (nz, ny, nx) array")
        if not hasattr(self, "volumes"):
            self.volumes = []
        self.volumes.append(dict(kind=kinds[kind], sigma_a=rgb(sigma_a), sigma_s=rgb(sigma_s), le=rgb(le), g=float(g), p0=tuple(p0),
                                 p1=tuple(p1), v2w=v2w, a=float(a), b=float(b), up=tuple(updir), density=d))
        return len(self.volumes) - 1

    def volume_integrator(self, kind="emission", stepsize=1.0):
        """VolumeIntegrator "emission" (the default, render_options.dart:24-39) | "single"; `stepsize` as in the scene file."""
        self.vol_integrator = ({"emission": 0, "single": 1}[kind], float(stepsize))

    def point_light(self, pos, intensity) -> int:
        self.lights.append(dict(kind=1, L=tuple(intensity), pos=tuple(pos), nsamples=1, shapes=[]))
        return len(self.lights) - 1

    def infinite_light(self, L=(1.0, 1.0, 1.0), nsamples=1, light_to_world=None, texels=None) -> int:
        """LightSource "infinite" (infinite_area_light.dart:37-69,308-316).  `texels`: level 0 of the light's radiance MIPMap as
        the reference holds it — (h, w, 3) float32 at power-of-two resolution, already multiplied by L (:50-53) — or None for
        the 1x1 white map of a scene without "mapname" (:66-68).  Radiance = lookup * L (:240-242)."""
        l2w = np.eye(4, dtype=np.float32) if light_to_world is None else np.asarray(light_to_world, np.float32).reshape(4, 4)
        tex = np.ones((1, 1, 3), np.float32) if texels is None else np.ascontiguousarray(texels, dtype=np.float32)
        self.lights.append(dict(kind=4, L=tuple(L), pos=(0, 0, 0), nsamples=nsamples, shapes=[], l2w=l2w, texels=tex))
        return len(self.lights) - 1

    def projection_light(self, I, fov=45.0, light_to_world=None, texels=None) -> int:
        """LightSource "projection" (projection_light.dart:38-100,141-150): a point light at the light's origin whose intensity is
        masked by a map projected along +z of the light's frame.  `texels`: level 0 of the map's MIPMap ((h, w, 3) float32,
        power of two) or None."""
        l2w = np.eye(4, dtype=np.float32) if light_to_world is None else np.asarray(light_to_world, np.float32).reshape(4, 4)
        pos = transform_points(l2w, [[0.0, 0.0, 0.0]])[0]
        aspect = 1.0 if texels is None else np.asarray(texels).shape[1] / np.asarray(texels).shape[0]
        screen = (-aspect, aspect, -1.0, 1.0) if aspect > 1.0 else (-1.0, 1.0, -1.0 / aspect, 1.0 / aspect)
        self.lights.append(dict(kind=5, L=tuple(I), pos=tuple(float(v) for v in pos), nsamples=1, shapes=[], w2l=mat_inv(l2w).reshape(16),
                                texels=None if texels is None else np.ascontiguousarray(texels, dtype=np.float32),
                                proj=perspective(fov, 1.0e-3, 1.0e30).reshape(16), screen=screen, hither=1.0e-3))
        return len(self.lights) - 1

    def goniometric_light(self, I, light_to_world=None, texels=None) -> int:
        """LightSource "goniometric" (goniometric_light.dart:37-86,117-123): a point light scaled by a lat-long map of directions
        (the light's y axis is the map's pole)."""
        l2w = np.eye(4, dtype=np.float32) if light_to_world is None else np.asarray(light_to_world, np.float32).reshape(4, 4)
        pos = transform_points(l2w, [[0.0, 0.0, 0.0]])[0]
        self.lights.append(dict(kind=6, L=tuple(I), pos=tuple(float(v) for v in pos), nsamples=1, shapes=[], w2l=mat_inv(l2w).reshape(16),
                                texels=None if texels is None else np.ascontiguousarray(texels, dtype=np.float32)))
        return len(self.lights) - 1

    def distant_light(self, frm, to, L) -> int:
        """DistantLight.Create (distant_light.dart:83-91) with an identity light-to-world: lightDir = normalize(from - to)."""
        d = (np.asarray(frm, np.float32).astype(np.float64) - np.asarray(to, np.float32).astype(np.float64)).astype(np.float32)
        d = (d.astype(np.float64) / math.sqrt(float((d.astype(np.float64) ** 2).sum()))).astype(np.float32)
        self.lights.append(dict(kind=2, L=tuple(L), pos=tuple(float(v) for v in d), nsamples=1, shapes=[]))
        return len(self.lights) - 1

    def spot_light(self, frm, to, I, coneangle=30.0, conedelta=5.0) -> int:
        """SpotLight.Create (spot_light.dart:88-118) with an identity CTM: light2world = Translate(from) * Inverse(dirToZ)."""
        frm64, to64 = np.asarray(frm, np.float64), np.asarray(to, np.float64)
        d = to64 - frm64
        d = (d / np.linalg.norm(d)).astype(np.float32).astype(np.float64)
        # Vector.CoordinateSystem (vector.dart:198-214)
        if abs(d[0]) > abs(d[1]):
            inv = 1.0 / math.sqrt(d[0] * d[0] + d[2] * d[2])
            du = np.array([-d[2] * inv, 0.0, d[0] * inv])
        else:
            inv = 1.0 / math.sqrt(d[1] * d[1] + d[2] * d[2])
            du = np.array([0.0, d[2] * inv, -d[1] * inv])
        du = du.astype(np.float32).astype(np.float64)
        dv = np.cross(d, du).astype(np.float32).astype(np.float64)
        dir_to_z = np.eye(4, dtype=np.float32)
        dir_to_z[0, :3], dir_to_z[1, :3], dir_to_z[2, :3] = du, dv, d
        l2w = mat_mul(translate(*[float(v) for v in np.asarray(frm, np.float32)]), mat_inv(dir_to_z))
        w2l = mat_inv(l2w)
        pos = transform_points(l2w, np.zeros((1, 3), np.float32))[0]
        self.lights.append(dict(kind=3, L=tuple(I), pos=tuple(float(v) for v in pos), nsamples=1, shapes=[],
                                w2l=_m(w2l).reshape(16), cos=(math.cos(math.radians(coneangle)), math.cos(math.radians(coneangle - conedelta)))))
        return len(self.lights) - 1

    def _area_light(self, L, nsamples):
        self.lights.append(dict(kind=0, L=tuple(L), pos=(0, 0, 0), nsamples=nsamples, shapes=[]))
        return len(self.lights) - 1

    def mesh(self, P, idx, material=0, o2w=None, area_light=None, nsamples=1, reverse=False, N=None, S=None, uv=None, alpha=None) -> int:
        """Shape "trianglemesh": vertices go to world space at construction (triangle_mesh.dart:24-37).
        `area_light` = emitted radiance: one DiffuseAreaLight per shape (dartray.dart:378-467).
        N / S / uv: the per-vertex "normal N" / "vector S" / "float uv" parameters (triangle_mesh.dart:88-140), kept in object
        space as the reference keeps them; Triangle.getShadingGeometry transforms them per hit (triangle.dart:271-364).
        alpha: the mesh's float "alpha" texture (triangle_mesh.dart:176-192).  Triangle.intersect / intersectP drop a hit only where
        it evaluates to exactly 0 (triangle.dart:139-151,196-237), so a constant-valued texture other than 0 changes nothing and is
        accepted; a mesh that is transparent everywhere, or an alpha that reads the hit point, is not on the GPU path."""
        if alpha is not None:
            a = fold_texture(alpha)
            if not _is_num(a):
                raise ValueError("alpha is a float texture (triangle_mesh.dart:176-192)")
            if a == 0.0:
                raise GpuUnsupported("a triangle mesh whose alpha texture is 0 everywhere")
        P = np.asarray(P, dtype=np.float32).reshape(-1, 3)
        nv = P.shape[0]
        m2w = np.eye(4, dtype=np.float32) if o2w is None else np.asarray(o2w, dtype=np.float32).reshape(4, 4)
        flags = (1 if N is not None else 0) | (2 if S is not None else 0) | (4 if uv is not None else 0)
        self.mesh_info.append((m2w, mat_inv(m2w), flags))
        self.vertN.append(np.zeros((nv, 3), np.float32) if N is None else np.asarray(N, np.float32).reshape(nv, 3))
        self.vertS.append(np.zeros((nv, 3), np.float32) if S is None else np.asarray(S, np.float32).reshape(nv, 3))
        self.vertUV.append(np.zeros((nv, 2), np.float32) if uv is None else np.asarray(uv, np.float32).reshape(nv, 2))
        self.tri_mesh += [len(self.mesh_info) - 1] * np.asarray(idx).reshape(-1, 3).shape[0]
        if o2w is not None:
            P = transform_points(o2w, P)
        idx = np.asarray(idx, dtype=np.uint32).reshape(-1, 3)
        first = len(self.tri_mat)
        light = -1
        if area_light is not None:
            light = self._area_light(area_light, nsamples)
            # ShapeSet refines LIFO (shape_set.dart:26-41): triangles in reverse order
            self.lights[light]["shapes"] = [("tri", first + k) for k in range(idx.shape[0] - 1, -1, -1)]
        self.P.append(P)
        self.idx.append(idx + self._nverts)
        self._nverts += P.shape[0]
        n = idx.shape[0]
        self.tri_mat += [material] * n
        self.tri_light += [light] * n
        self.tri_rev += [1 if reverse else 0] * n
        self._order.append(("mesh", first, n))
        return first

    def sphere(self, o2w, radius=1.0, zmin=None, zmax=None, phimax=360.0, material=0, area_light=None, nsamples=1,
               reverse=False) -> int:
        o2w = np.asarray(o2w, dtype=np.float32).reshape(4, 4)
        light = -1
        k = len(self.sph)
        if area_light is not None:
            light = self._area_light(area_light, nsamples)
            self.lights[light]["shapes"] = [("sph", k)]
        self.sph.append((o2w, mat_inv(o2w), (radius, -radius if zmin is None else zmin, radius if zmax is None else zmax, phimax),
                         material, light, 1 if reverse else 0))
        self._order.append(("sphere", k))
        return k

    def disk(self, o2w, height=0.0, radius=1.0, innerradius=0.0, phimax=360.0, material=0, area_light=None, nsamples=1,
             reverse=False) -> int:
        """Shape "disk" (lib/shapes/disk.dart:157-166)."""
        o2w = np.asarray(o2w, dtype=np.float32).reshape(4, 4)
        light = -1
        k = len(self.dsk)
        if area_light is not None:
            light = self._area_light(area_light, nsamples)
            self.lights[light]["shapes"] = [("dsk", k)]
        self.dsk.append((o2w, mat_inv(o2w), (height, radius, innerradius, phimax), material, light, 1 if reverse else 0))
        self._order.append(("disk", k))
        return k

    def _quadric(self, kind, o2w, params, material, area_light, nsamples, reverse) -> int:
        o2w = np.asarray(o2w, dtype=np.float32).reshape(4, 4)
        light = -1
        k = len(self.quad)
        if area_light is not None:
            if kind != QUADRIC_CYLINDER:  # Shape.sample is unimplemented for them (lib/core/shape.dart:83-86)
                raise ValueError("only the cylinder among these quadrics can be an area light (it alone has sample())")
            light = self._area_light(area_light, nsamples)
            self.lights[light]["shapes"] = [("quad", k)]
        prm = tuple(params) + (0.0,) * (8 - len(params))
        self.quad.append((kind, o2w, mat_inv(o2w), prm, material, light, 1 if reverse else 0))
        self._order.append(("quad", k))
        return k

    def cylinder(self, o2w, radius=1.0, zmin=-1.0, zmax=1.0, phimax=360.0, material=0, area_light=None, nsamples=1,
                 reverse=False) -> int:
        """Shape "cylinder" (lib/shapes/cylinder.dart:239-247)."""
        return self._quadric(QUADRIC_CYLINDER, o2w, (radius, zmin, zmax, phimax), material, area_light, nsamples, reverse)

    def cone(self, o2w, radius=1.0, height=1.0, phimax=360.0, material=0, reverse=False) -> int:
        """Shape "cone" (lib/shapes/cone.dart:216-222)."""
        return self._quadric(QUADRIC_CONE, o2w, (height, radius, phimax), material, None, 1, reverse)

    def paraboloid(self, o2w, radius=1.0, zmin=0.0, zmax=1.0, phimax=360.0, material=0, reverse=False) -> int:
        """Shape "paraboloid" (lib/shapes/paraboloid.dart:220-228)."""
        return self._quadric(QUADRIC_PARABOLOID, o2w, (radius, zmin, zmax, phimax), material, None, 1, reverse)

    def hyperboloid(self, o2w, p1=(0.0, 0.0, 0.0), p2=(1.0, 1.0, 1.0), phimax=360.0, material=0, reverse=False) -> int:
        """Shape "hyperboloid" (lib/shapes/hyperboloid.dart:263-268)."""
        return self._quadric(QUADRIC_HYPERBOLOID, o2w, tuple(p1) + tuple(p2) + (phimax,), material, None, 1, reverse)

    def arrays(self) -> dict:
        ntris = len(self.tri_mat)
        nsph = len(self.sph)
        ndsk = len(self.dsk)
        P = np.concatenate(self.P) if self.P else np.zeros((0, 3), np.float32)
        idx = np.concatenate(self.idx) if self.idx else np.zeros((0, 3), np.uint32)
        # refined order handed to BVHAccel: Primitive.fullyRefine is LIFO per primitive (primitive.dart:71-84)
        if self._order is not self._top_order:
            raise ValueError("ObjectBegin without ObjectEnd")
        nprims = ntris + nsph + ndsk + len(self.quad)

        def refine(items):
            order = []
            for item in items:
                if item[0] == "mesh":
                    order += list(range(item[1] + item[2] - 1, item[1] - 1, -1))
                elif item[0] == "sphere":
                    order.append(ntris + item[1])
                elif item[0] == "disk":
                    order.append(ntris + nsph + item[1])
                elif item[0] == "instance":
                    order.append(nprims + item[1])
                else:
                    order.append(ntris + nsph + ndsk + item[1])
            return order
        order = refine(self._top_order)
        object_orders = [refine(o[0]) for o in self.objects]
        lights = []
        base = {"tri": 0, "sph": ntris, "dsk": ntris + nsph, "quad": ntris + nsph + ndsk}
        for l in self.lights:
            shapes = [base[s[0]] + s[1] for s in l["shapes"]]
            lights.append(dict(kind=l["kind"], L=l["L"], pos=l["pos"], nsamples=l["nsamples"], shapes=shapes,
                               l2w=l.get("l2w"), texels=l.get("texels"), proj=l.get("proj"), screen=l.get("screen"), hither=l.get("hither", 1.0e-3),
                               w2l=l.get("w2l", np.eye(4, dtype=np.float32).reshape(16)), cos=l.get("cos", (0.0, 0.0))))
        mats = self.materials or [(0, (0.5, 0.5, 0.5), 0.0)]
        general = any(m[0] in ("lobes", "program") for m in mats)
        lobe_lists = [m[1] if m[0] == "lobes" else ([] if m[0] == "program" else matte_lobes(m[1], m[2])) for m in mats]
        table, programs = TextureTable(), np.zeros(len(mats), PROG_DTYPE)
        programs["kind"], programs["tex"], programs["bump"], programs["m1"], programs["m2"] = -1, -1, -1, -1, -1
        for i, m in enumerate(mats):
            if m[0] != "program":
                continue
            programs[i]["kind"] = m[1]
            for j, (v, spectrum) in enumerate(m[2]):
                programs[i]["tex"][j] = table.add(v, spectrum)
            if m[3] is not None:
                programs[i]["bump"] = table.add(m[3], False)
            if m[1] == 11:
                programs[i]["m1"], programs[i]["m2"] = m[4], 0
            if m[1] == 9:
                programs[i]["m1"], programs[i]["m2"] = m[4], m[5]
                for sub in (m[4], m[5]):
                    if mats[sub][0] == "program" and mats[sub][1] == 9:
                        raise ValueError("a mix of mixes (ScaledBxDF of a ScaledBxDF) is not representable")
        tex_nodes, tex_texels = table.arrays()
        has_programs = any(m[0] == "program" for m in mats)
        lobes = [l for ll in lobe_lists for l in ll]
        if general:  # the matte-only entry point cannot carry these: upload_scene uses drt_set_material_lobes instead
            mats = [(0, (0.0, 0.0, 0.0), 0.0) for _ in mats]
        return dict(
            P=P, idx=idx, tri_mat=np.asarray(self.tri_mat, np.int32), tri_light=np.asarray(self.tri_light, np.int32),
            tri_rev=np.asarray(self.tri_rev, np.uint8),
            sph_o2w=np.stack([s[0].reshape(16) for s in self.sph]) if self.sph else np.zeros((0, 16), np.float32),
            sph_w2o=np.stack([s[1].reshape(16) for s in self.sph]) if self.sph else np.zeros((0, 16), np.float32),
            sph_params=np.asarray([s[2] for s in self.sph], np.float64).reshape(-1, 4),
            sph_mat=np.asarray([s[3] for s in self.sph], np.int32), sph_light=np.asarray([s[4] for s in self.sph], np.int32),
            sph_rev=np.asarray([s[5] for s in self.sph], np.uint8),
            dsk_o2w=np.stack([s[0].reshape(16) for s in self.dsk]) if self.dsk else np.zeros((0, 16), np.float32),
            dsk_w2o=np.stack([s[1].reshape(16) for s in self.dsk]) if self.dsk else np.zeros((0, 16), np.float32),
            dsk_params=np.asarray([s[2] for s in self.dsk], np.float64).reshape(-1, 4),
            dsk_mat=np.asarray([s[3] for s in self.dsk], np.int32), dsk_light=np.asarray([s[4] for s in self.dsk], np.int32),
            dsk_rev=np.asarray([s[5] for s in self.dsk], np.uint8),
            mesh_shading=any(m[2] for m in self.mesh_info),
            mesh_of_tri=np.asarray(self.tri_mesh, np.uint32),
            mesh_o2w=np.stack([m[0].reshape(16) for m in self.mesh_info]) if self.mesh_info else np.zeros((0, 16), np.float32),
            mesh_w2o=np.stack([m[1].reshape(16) for m in self.mesh_info]) if self.mesh_info else np.zeros((0, 16), np.float32),
            mesh_flags=np.asarray([m[2] for m in self.mesh_info], np.uint8),
            vert_N=np.concatenate(self.vertN) if self.vertN else np.zeros((0, 3), np.float32),
            vert_S=np.concatenate(self.vertS) if self.vertS else np.zeros((0, 3), np.float32),
            vert_uv=np.concatenate(self.vertUV) if self.vertUV else np.zeros((0, 2), np.float32),
            quad_kind=np.asarray([q[0] for q in self.quad], np.int32),
            quad_o2w=np.stack([q[1].reshape(16) for q in self.quad]) if self.quad else np.zeros((0, 16), np.float32),
            quad_w2o=np.stack([q[2].reshape(16) for q in self.quad]) if self.quad else np.zeros((0, 16), np.float32),
            quad_params=np.asarray([q[3] for q in self.quad], np.float64).reshape(-1, 8),
            quad_mat=np.asarray([q[4] for q in self.quad], np.int32), quad_light=np.asarray([q[5] for q in self.quad], np.int32),
            quad_rev=np.asarray([q[6] for q in self.quad], np.uint8),
            order=np.asarray(order, np.uint32),
            mat_kind=np.asarray([m[0] for m in mats], np.int32), mat_kd=np.asarray([m[1] for m in mats], np.float32),
            mat_sigma=np.asarray([m[2] for m in mats], np.float32),
            mat_general=general,
            mat_programs=programs if has_programs else None, tex_nodes=tex_nodes, tex_texels=tex_texels,
            measured=list(getattr(self, "measured", [])),
            object_offsets=np.asarray(np.cumsum([0] + [len(o) for o in object_orders]), np.uint32),
            object_prims=np.asarray([p for o in object_orders for p in o], np.uint32),
            object_split=np.asarray([o[1] for o in self.objects], np.int32),
            object_max_node_prims=np.asarray([o[2] for o in self.objects], np.int32),
            instance_object=np.asarray([i[0] for i in self.instances], np.uint32),
            instance_start_m=np.asarray([i[1] for i in self.instances], np.float32).reshape(-1, 16),
            instance_start_minv=np.asarray([i[2] for i in self.instances], np.float32).reshape(-1, 16),
            instance_end_m=np.asarray([i[3] for i in self.instances], np.float32).reshape(-1, 16),
            instance_end_minv=np.asarray([i[4] for i in self.instances], np.float32).reshape(-1, 16),
            instance_times=np.asarray([(i[5], i[6]) for i in self.instances], np.float64).reshape(-1, 2),
            mat_lobe_offsets=np.asarray(np.cumsum([0] + [len(ll) for ll in lobe_lists]), np.uint32),
            lobe_kind=np.asarray([l["kind"] for l in lobes], np.int32),
            lobe_rgb=np.asarray([l["rgb"] for l in lobes], np.float32).reshape(-1, 3),
            lobe_fresnel=np.asarray([l["fresnel"] for l in lobes], np.int32),
            lobe_eta=np.asarray([l["eta"] for l in lobes], np.float32).reshape(-1, 3),
            lobe_k=np.asarray([l["k"] for l in lobes], np.float32).reshape(-1, 3),
            lobe_scalars=np.asarray([(l["param"], l["ei"], l["et"]) for l in lobes], np.float64).reshape(-1, 3),
            lobe_wrap=np.asarray([l.get("wrap", 0) for l in lobes], np.int32),
            lobe_scale=np.asarray([l.get("scale", (1, 1, 1)) for l in lobes], np.float32).reshape(-1, 3),
            light_kind=np.asarray([l["kind"] for l in lights], np.int32),
            light_L=np.asarray([l["L"] for l in lights], np.float32).reshape(-1, 3),
            light_pos=np.asarray([l["pos"] for l in lights], np.float32).reshape(-1, 3),
            light_nsamples=np.asarray([l["nsamples"] for l in lights], np.int32),
            light_w2l=np.asarray([l["w2l"] for l in lights], np.float32).reshape(-1, 16),
            light_cos=np.asarray([l["cos"] for l in lights], np.float64).reshape(-1, 2),
            light_has_spot=any(l["kind"] == 3 for l in lights),
            light_infinite=[(i, l["l2w"], l["texels"]) for i, l in enumerate(lights) if l["kind"] == 4],
            light_mapped=[(i, l["texels"], l["w2l"], l["proj"], l["screen"], l["hither"]) for i, l in enumerate(lights) if l["kind"] >= 5],
            light_shape_offsets=np.asarray(np.cumsum([0] + [len(l["shapes"]) for l in lights]), np.uint32),
            light_shape_prims=np.asarray([p for l in lights for p in l["shapes"]], np.uint32),
            volumes=list(getattr(self, "volumes", [])),
            vol_integrator=getattr(self, "vol_integrator", (0, 1.0)),
        )


def pack_volumes(volumes: list) -> dict:
    """The flat arrays of drt_set_volumes for a list of SceneBuilder.volume() entries."""
    n = len(volumes)
    f32 = lambda rows, w: np.ascontiguousarray(np.asarray(rows, np.float32).reshape(n, w)) if n else np.zeros((0, w), np.float32)
    dens, off = [], [0]
    for v in volumes:
        if v["density"] is not None:
            dens.append(v["density"].reshape(-1))
        off.append(off[-1] + (v["density"].size if v["density"] is not None else 0))
    return dict(
        n=n, kind=np.asarray([v["kind"] for v in volumes], np.int32),
        sigma_a=f32([v["sigma_a"] for v in volumes], 3), sigma_s=f32([v["sigma_s"] for v in volumes], 3), le=f32([v["le"] for v in volumes], 3),
        g=np.asarray([v["g"] for v in volumes], np.float64),
        p0p1=f32([tuple(v["p0"]) + tuple(v["p1"]) for v in volumes], 6),
        v2w=f32([v["v2w"].reshape(16) for v in volumes], 16), w2v=f32([mat_inv(v["v2w"]).reshape(16) for v in volumes], 16),
        ab=np.ascontiguousarray(np.asarray([(v["a"], v["b"]) for v in volumes], np.float64).reshape(n, 2)),
        up=f32([v["up"] for v in volumes], 3),
        dims=np.ascontiguousarray(np.asarray([(v["density"].shape[2], v["density"].shape[1], v["density"].shape[0]) if v["density"] is not None
                                              else (1, 1, 1) for v in volumes], np.int32).reshape(n, 3)),
        density_offsets=np.asarray(off, np.uint64),
        density=np.concatenate(dens).astype(np.float64) if dens else np.zeros(1, np.float64))


def upload_scene(ctx, arrays: dict, split: int = 2, max_node_prims: int = 4):
    """Drives either dartray_b200.capi.Context or tests.oracle_lib.Oracle (same method names)."""
    a = arrays
    ctx.set_triangles(a["P"], a["idx"], a["tri_mat"], a["tri_light"], a["tri_rev"])
    if a.get("mesh_shading"):
        fl = a["mesh_flags"]
        ctx.set_mesh_shading(a["vert_N"] if (fl & 1).any() else None, a["vert_S"] if (fl & 2).any() else None,
                             a["vert_uv"] if (fl & 4).any() else None, a["mesh_of_tri"], a["mesh_o2w"], a["mesh_w2o"], fl)
    ctx.set_spheres(a["sph_o2w"], a["sph_w2o"], a["sph_params"], a["sph_mat"], a["sph_light"], a["sph_rev"])
    if a["dsk_params"].shape[0]:
        ctx.set_disks(a["dsk_o2w"], a["dsk_w2o"], a["dsk_params"], a["dsk_mat"], a["dsk_light"], a["dsk_rev"])
    qk = a.get("quad_kind", np.zeros(0, np.int32))
    i = 0
    while i < qk.shape[0]:  # one call per run of equal kinds keeps the ids in append order
        j = i
        while j < qk.shape[0] and qk[j] == qk[i]:
            j += 1
        ctx.set_quadrics(int(qk[i]), a["quad_o2w"][i:j], a["quad_w2o"][i:j], a["quad_params"][i:j], a["quad_mat"][i:j],
                         a["quad_light"][i:j], a["quad_rev"][i:j])
        i = j
    if a.get("instance_object") is not None and a["instance_object"].shape[0]:
        ctx.set_instances(a["object_offsets"], a["object_prims"], a["object_split"], a["object_max_node_prims"], a["instance_object"],
                          a["instance_start_m"], a["instance_start_minv"], a["instance_end_m"], a["instance_end_minv"], a["instance_times"])
    ctx.set_build_order(a["order"])
    ctx.build_bvh(split, max_node_prims)
    if a.get("measured"):
        ctx.set_measured(a["measured"])
    if a.get("mat_general"):
        ctx.set_material_lobes(a["mat_lobe_offsets"], a["lobe_kind"], a["lobe_rgb"], a["lobe_fresnel"], a["lobe_eta"], a["lobe_k"],
                               a["lobe_scalars"])
        if "lobe_wrap" in a and a["lobe_wrap"].any():
            ctx.set_lobe_wrappers(a["lobe_wrap"], a["lobe_scale"])
    else:
        ctx.set_materials(a["mat_kind"], a["mat_kd"], a["mat_sigma"])
    if a.get("mat_programs") is not None:
        ctx.set_textures(a["tex_nodes"], a["tex_texels"])
        ctx.set_material_programs(a["mat_programs"])
    ctx.set_lights(a["light_kind"], a["light_L"], a["light_pos"], a["light_nsamples"], a["light_shape_offsets"],
                   a["light_shape_prims"])
    if a.get("light_has_spot"):
        ctx.set_spot_params(a["light_w2l"], a["light_cos"])
    for i, l2w, tex in a.get("light_infinite", []):
        ctx.set_infinite_light(i, tex, l2w, mat_inv(l2w))
    for i, tex, w2l, proj, screen, hither in a.get("light_mapped", []):
        ctx.set_light_map(i, tex, w2l, proj, screen, hither)
    ctx.set_volumes(pack_volumes(a.get("volumes") or []))
    ctx.set_volume_integrator(*a.get("vol_integrator", (0, 1.0)))


def configure_render(ctx, camera: PerspectiveCamera, film: Film, sampler: Sampler, integrator: Integrator):
    ctx.set_camera(camera.raster_to_camera(film.xres, film.yres), camera.camera_to_world, camera.lens_radius,
                   camera.focal_distance, camera.shutter_open, camera.shutter_close)
    # an animated camera: the end-time CTM and the two transform times (Camera.cameraToWorld is an AnimatedTransform, camera.dart:27)
    end = getattr(camera, "camera_to_world_end", None)
    ctx.set_camera_motion(end, *getattr(camera, "transform_times", (0.0, 1.0)))
    ctx.set_camera_kind(getattr(camera, "kind", 0))
    xw, yw, table = film.table()
    ctx.set_film(film.xres, film.yres, film.crop, xw, yw, table)
    ctx.set_sampler(sampler.kind, sampler.xs, sampler.ys, sampler.spp, int(sampler.jitter), sampler.pixel_order,
                    sampler.tile_size, sampler.seed, sampler.rng_mode)
    if getattr(sampler, "sample_table", None) is not None:
        ctx.set_sample_table(sampler.sample_table)
    ctx.set_integrator(integrator.kind, integrator.maxdepth, integrator.strategy, integrator.ao_nsamples,
                       integrator.ao_mindist, integrator.ao_maxdist)
