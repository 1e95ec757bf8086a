"""blenderraytracer_b200 — B200-native (sm_100a) drop-in for the render path of Shinzef/BlenderRayTracer.

Only what the hot path needs lives here: ``csrc/`` (hand-written CUDA kernels + the C ABI of include/brt.h, built
into ``libbrt.so``) and the host-side mirror of the reference's ``RayTracer`` / scene classes.  Importing the package
does not touch the GPU; constructing a ``RayTracer`` does, and fails loudly without one (no CPU fallback).
"""
from ._lib import BrtError, LIB_PATH, load  # noqa: F401
from .raytracer import RayTracer, make_perm  # noqa: F401
from .scene import (Box, CheckerTexture, Dielectric, DirectionalLight, Emissive, Lambertian, MarbleTexture, Metal,  # noqa: F401
                    NoiseTexture, Plane, PointLight, SolidColor, Sphere, TexturedLambertian, TexturedMetal, Triangle,
                    TriangleMesh, WoodTexture, World)

__all__ = ["RayTracer", "World", "Sphere", "Plane", "Box", "Triangle", "TriangleMesh", "Lambertian", "Metal", "Dielectric",
           "Emissive", "PointLight", "DirectionalLight", "SolidColor", "CheckerTexture", "NoiseTexture", "MarbleTexture", "WoodTexture",
           "TexturedLambertian", "TexturedMetal", "BrtError", "make_perm", "load", "LIB_PATH"]
