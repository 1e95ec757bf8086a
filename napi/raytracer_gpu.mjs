// raytracer_gpu.mjs — drop-in GPU render for the reference's RayTracer (js/ray-tracer.js).
//
//   import { RayTracer } from '../js/ray-tracer.js';
//   import { installGpuRender } from './napi/raytracer_gpu.mjs';
//   installGpuRender(RayTracer);            // RayTracer.prototype.render now runs on the B200 through libbrt
//
// Everything above render() stays the reference's own code: scene-loader.js builds the World, camera.js builds the
// Camera, the UI setters mutate the settings.  This shim only FLATTENS those live objects (world.js:9-18,
// geometry.js, materials.js, lights.js, camera.js:14-35) into the typed arrays brt_addon.c takes and keeps the
// observable contract of render(onProgress) (ray-tracer.js:166-281): imageData.data is filled (row 0 = top, alpha 255),
// putImageData is called, onProgress(fraction) fires and ends with 1.0, window.renderCancelled stops the render without
// the final blit.  The image is written by the GPU directly into imageData.data's backing store.
//
// NOTE: the build image has no Node.js.  This file is executed there, unmodified, by baseline/minijs.py (the interpreter that also
// runs the reference's js/*.js) over napi/napi_host.py, a Node-API host that loads the REAL brt_addon.node: shim -> addon ->
// libbrt, with brt_render on a worker thread, progress through a thread-safe function and the cancel poll below ticking while
// it runs (tests/test_js_shim.py: flattening checked against the reference's live objects; image equal to the Python ctypes
// binding's, byte for byte, on the GPU).  brt_addon.node alone is also driven by napi/mock_node_host.c (tests/test_napi_mock.py).
//
//   installGpuRender(RayTracer, { devices: [0, 1, 2, 3, 4, 5, 6, 7] })    // one ctx over 8 GPUs: render() is still ONE call
import { createRequire } from 'node:module';
const require = createRequire(import.meta.url);
const addon = require('./brt_addon.node');

const OBJ = { Sphere: 0, Plane: 1, Box: 2, Triangle: 3, TriangleMesh: 4 };
const MAT = { Lambertian: 0, Metal: 1, Dielectric: 2, Emissive: 3, TexturedLambertian: 0, TexturedMetal: 1 };
const TEX = { SolidColor: 0, CheckerTexture: 1, NoiseTexture: 2, MarbleTexture: 3, WoodTexture: 4 };
const AA = { none: 0, supersampling: 1, stochastic: 2 };            // any other string: pixel-centre samples (3)
const TONEMAP = { reinhard: 0, aces: 1, linear: 2 };
const CAM = { perspective: 0, orthographic: 1 };                     // any other string: 2 (camera.js:25 vs :39)

const v3 = (v) => [v.x, v.y, v.z];

// world.objects / world.lights -> rows (one per object, IN ORDER: the row index is the object ID, world.js:24-30)
export function flattenWorld(world) {
  const objects = [], materials = [], tris = [], lights = [], textures = [], perms = [];
  let firstTri = 0;
  for (const o of world.objects) {
    const kind = o.constructor.name;
    // a TriangleMesh has no `material` member: the constructor hands it to its triangles (geometry.js:231); a mesh whose every
    // triangle was filtered out can never be hit, so any material does
    const one = { x: 1, y: 1, z: 1 };
    const m = o.material ?? (kind === 'TriangleMesh' && o.triangles.length ? o.triangles[0].material : undefined) ?? { albedo: one, constructor: { name: 'Lambertian' } };
    const mk = m.constructor.name;
    const color = mk === 'Emissive' ? m.color : (m.albedo ?? one);
    const param = (mk === 'Metal' || mk === 'TexturedMetal') ? m.roughness : mk === 'Dielectric' ? m.refractionIndex : mk === 'Emissive' ? m.intensity : 0;
    if (!(mk in MAT)) throw new Error(`unsupported material ${mk}`);
    let tex = 0;
    if (m.texture) {                                                  // TexturedLambertian / TexturedMetal (materials.js:99-126)
      const t = m.texture, tk = t.constructor.name;
      if (!(tk in TEX)) throw new Error(`unsupported texture ${tk}`);
      textures.push(TEX[tk], ...v3(t.odd ?? t.color ?? one), ...v3(t.even ?? one), t.scale ?? 1);
      perms.push(...(t.noise ? t.noise.p.slice(0, 256) : Array.from({ length: 256 }, (_, i) => i)));
      tex = textures.length / 8;                                     // 1-based
    }
    materials.push(MAT[mk], ...v3(color), param, tex);
    const mat = materials.length / 6 - 1;
    const z = [0, 0, 0];
    if (kind === 'Sphere') objects.push(OBJ.Sphere, mat, ...v3(o.center), o.radius, 0, 0, ...z, 0, 0);
    else if (kind === 'Plane') objects.push(OBJ.Plane, mat, ...v3(o.point), ...v3(o.normal), ...z, 0, 0);
    else if (kind === 'Box') objects.push(OBJ.Box, mat, ...v3(o.min), ...v3(o.max), ...z, 0, 0);
    else if (kind === 'Triangle') objects.push(OBJ.Triangle, mat, ...v3(o.v0), ...v3(o.v1), ...v3(o.v2), 0, 0);
    else if (kind === 'TriangleMesh') {
      // o.triangles is already post-filter (geometry.js:206-231): its order defines the triangle IDs
      for (const t of o.triangles) tris.push(...v3(t.v0), ...v3(t.v1), ...v3(t.v2));
      objects.push(OBJ.TriangleMesh, mat, ...z, ...z, ...z, firstTri, o.triangles.length);
      firstTri += o.triangles.length;
    } else throw new Error(`unsupported object ${kind}`);
  }
  for (const l of world.lights) {
    const point = l.constructor.name === 'PointLight';
    lights.push(point ? 0 : 1, ...v3(point ? l.position : l.direction), ...v3(l.color), l.intensity);
  }
  return { objects: new Float64Array(objects), materials: new Float64Array(materials), tris: new Float64Array(tris), lights: new Float64Array(lights),
    textures: new Float64Array(textures), perms: Uint8Array.from(perms) };
}

// the Camera object's own derived members (camera.js:14-35): no libm call is repeated on the native side
export function flattenCamera(c) {
  return new Float64Array([...v3(c.origin), ...v3(c.lowerLeftCorner), ...v3(c.horizontal), ...v3(c.vertical),
    ...v3(c.u), ...v3(c.v), ...v3(c.w), c.lensRadius, c.type in CAM ? CAM[c.type] : 2]);
}

// world.background is a closure and cannot cross a C ABI: the wrappers below remember which factory installed it;
// a bound skyGradient / proceduralSky can still be recognised by name.
function backgroundOf(rt) {
  const w = rt.world, bg = w.background;
  if (rt._brtBackground) return rt._brtBackground;                   // set by the updateBackground wrapper
  if (bg === w.skyGradient || (bg && bg.name === 'bound skyGradient')) return { kind: 0, color: [0.1, 0.1, 0.1] };
  if (bg && bg.name === 'bound proceduralSky') return { kind: 3, color: [0.1, 0.1, 0.1] };
  return { kind: 0, color: [0.1, 0.1, 0.1] };
}

// options: device | devices (one ctx over several GPUs), seed, preview (blit at every progress callback), sppBatch (samples per
// progress step; 0 = libbrt decides)
export function installGpuRender(RayTracer, { device = 0, devices = undefined, seed = 1, preview = true, sppBatch = 0 } = {}) {
  const origUpdateBackground = RayTracer.prototype.updateBackground;
  RayTracer.prototype.updateBackground = function (type, intensity = 1.0) {       // ray-tracer.js:568-585
    origUpdateBackground.call(this, type, intensity);
    const kind = { solid: 1, hdri: 2, procedural_sky: 3 }[type] ?? 0;
    this._brtBackground = { kind, color: [0.1, 0.1, 0.1] };
  };
  const origPreset = RayTracer.prototype.loadPreset;
  RayTracer.prototype.loadPreset = function (name) {                              // ray-tracer.js:282-299: new World() => gradient
    origPreset.call(this, name);
    this._brtBackground = name === 'cornell' ? { kind: 1, color: [0, 0, 0] } : undefined;   // setupCornellBox (:424)
  };
  const origLoad = RayTracer.prototype.loadFromJSON;
  RayTracer.prototype.loadFromJSON = function (json) {                            // ray-tracer.js:305-334
    const ok = origLoad.call(this, json);
    if (!ok) return ok;                                              // a failed load leaves world, camera and background as they were (:329-332)
    const t = json?.background?.type;
    // deviation D1: the loader binds the solid / hdri FACTORIES (scene-loader.js:43,45 -> NaN -> black); we honour the intent
    this._brtBackground = { kind: { solid: 1, hdri: 2, procedural_sky: 3 }[t] ?? 0, color: json?.background?.color ?? [0.1, 0.1, 0.1] };
    return ok;
  };

  RayTracer.prototype.render = async function (onProgress) {                       // ray-tracer.js:166-281
    this._brt ??= addon.create(devices ?? device);                  // an array: brt_create_multi, samples split over the GPUs
    const ctx = this._brt;
    const flat = flattenWorld(this.world);
    // 1 = BRT_SCENE_CONSTRUCTED: these are members of constructed objects (Plane.normal and DirectionalLight.direction are already
    // normalised, Metal.roughness already clamped): libbrt must store them as they are
    addon.setSceneFlat(ctx, flat.objects, flat.materials, flat.tris, flat.lights, flat.textures, flat.perms, 1);
    addon.setCameraDerived(ctx, flattenCamera(this.camera));
    const bg = backgroundOf(this);
    const perm = this.world.cloudNoise ? Uint8Array.from(this.world.cloudNoise.p.slice(0, 256)) : undefined;
    addon.setBackground(ctx, bg.kind, bg.color[0], bg.color[1], bg.color[2], this.world.skyIntensity, perm);
    addon.setRenderParams(ctx, {
      width: this.width, height: this.height, samples: this.samples, maxBounces: this.maxBounces,
      aaMode: this.antiAliasing in AA ? AA[this.antiAliasing] : 3, toneMapping: TONEMAP[this.toneMapping] ?? 0,
      exposure: this.exposure, gamma: this.gamma, denoising: this.denoising ? 1 : 0, denoiseStrength: this.denoiseStrength,
      seed: seed + (this._brtFrame = (this._brtFrame ?? 0) + 1),     // Math.random is unseeded: a new stream per render
      // progressive preview: before every progress callback libbrt resolves the image of the samples traced so far into
      // imageData.data, and the callback below blits it — the reference blits finished rows as it goes (:236-238)
      preview: preview ? 1 : 0, sppBatch,
    });
    const poll = setInterval(() => { if (globalThis.window?.renderCancelled) addon.cancel(ctx); }, 50);   // :190,:256
    try {
      await addon.render(ctx, this.imageData.data, (fraction) => {
        if (preview && fraction < 1) this.ctx.putImageData(this.imageData, 0, 0);   // :236-238 (whole image of fewer samples instead of finished rows)
        if (onProgress) onProgress(fraction);                                        // :258-259
      });
    } catch (e) {
      if (e.code === 'BRT_E_CANCELLED') return;                      // the reference stops without the final blit (:264)
      throw e;
    } finally {
      clearInterval(poll);
    }
    this.ctx.putImageData(this.imageData, 0, 0);                     // :278
  };
  return RayTracer;
}
