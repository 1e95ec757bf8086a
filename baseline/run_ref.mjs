// run_ref.mjs — runs the UNMODIFIED reference renderer (Shinzef/BlenderRayTracer, js/*.js) under Node.js.
//
// The reference is browser JavaScript: RayTracer needs a canvas and `window`.  This harness supplies a fake canvas
// ({width, height, getContext: () => ({createImageData, putImageData})}), `globalThis.window = {renderCancelled: false}`, a
// silenced console, and calls RayTracer.render() (js/ray-tracer.js:166-281) exactly as js/ui-controller.js:189 does.  Nothing of
// the reference is edited or re-implemented here.  Two uses:
//
//  1. PINNING THE ORACLE (`--seed N`): Math.random (js/math.js:21-31, js/materials.js:62) is replaced by the same
//     Philox4x32-10 stream the oracle uses — counter (pixel, sample, block, 'BRT1'), key = seed, uniforms = top 24 bits / 2^24,
//     restarted at every getAntiAliasSample(i, j, s) call (ray-tracer.js:203), pixel = (H-1-j)*W + i — so the reference draws
//     the very numbers oracle/brt_oracle.cpp draws, in its own order.  The per-pixel mean radiance (argument of toneMap,
//     :209), the tone-mapped + gamma'd colour (result of gammaCorrect, :210) and imageData.data are captured and written as
//     JSON.  baseline/make_fixtures.mjs drives this over the 13 cases of tests/golden/reference_cases.json.
//  2. TIMING (`--time-rows N` / `--as-shipped`): native Math.random, wall clock.  `--as-shipped` times render() itself
//     (per-row console.log and setTimeout yields included, ray-tracer.js:192,261); `--time-rows N` times a tight loop over
//     getAntiAliasSample -> camera.getRay -> rayColor on N rows spread evenly over the frame (BASELINE.md §3 rows 1 and 2).
//     Prints one JSON line {msamples_per_s, sample, ...} that bench.py reads.
//
//   node baseline/run_ref.mjs --ref baseline/_ref/js --scene tests/golden/sample_scene.json --width 600 --height 400 \
//        --samples 16 --bounces 10 --time-rows 16
//   (stage the reference once: node baseline/make_fixtures.mjs --stage /path/to/BlenderRayTracer)
import fs from 'node:fs';
import path from 'node:path';
import { pathToFileURL } from 'node:url';

// ------------------------------------------------------------------------------------------------ Philox4x32-10 (Random123)
const M0 = 0xD2511F53n, M1 = 0xCD9E8D57n, W0 = 0x9E3779B9, W1 = 0xBB67AE85, TAG = 0x42525431;
export function philoxBlock(pixel, sample, block, seedLo, seedHi) {
  let c0 = pixel >>> 0, c1 = sample >>> 0, c2 = block >>> 0, c3 = TAG, k0 = seedLo >>> 0, k1 = seedHi >>> 0;
  for (let r = 0; r < 10; r++) {
    const p0 = M0 * BigInt(c0), p1 = M1 * BigInt(c2);
    const n0 = (Number(p1 >> 32n) ^ c1 ^ k0) >>> 0, n1 = Number(p1 & 0xFFFFFFFFn);
    const n2 = (Number(p0 >> 32n) ^ c3 ^ k1) >>> 0, n3 = Number(p0 & 0xFFFFFFFFn);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 = (k0 + W0) >>> 0; k1 = (k1 + W1) >>> 0;
  }
  return [c0, c1, c2, c3];
}
class PhiloxStream {
  constructor(seed) { const s = BigInt(seed); this.lo = Number(s & 0xFFFFFFFFn); this.hi = Number((s >> 32n) & 0xFFFFFFFFn); this.restart(0, 0); }
  restart(pixel, sample) { this.pixel = pixel; this.sample = sample; this.block = 0; this.buf = []; }
  next() {
    if (this.buf.length === 0) this.buf = philoxBlock(this.pixel, this.sample, this.block++, this.lo, this.hi);
    return (this.buf.shift() >>> 8) / 16777216;
  }
}

// ------------------------------------------------------------------------------------------------ fake browser
function fakeCanvas(width, height) {
  const ctx = {
    createImageData: (w, h) => ({ width: w, height: h, data: new Uint8ClampedArray(w * h * 4) }),
    putImageData() {},
  };
  return { width, height, style: {}, getContext: () => ctx };
}

export async function loadReference(refJsDir) {
  globalThis.window = globalThis.window ?? { renderCancelled: false };
  const url = pathToFileURL(path.resolve(refJsDir, 'ray-tracer.js')).href;
  const { RayTracer } = await import(url);
  const { Vec3 } = await import(pathToFileURL(path.resolve(refJsDir, 'math.js')).href);
  return { RayTracer, Vec3 };
}

function quiet(fn) {
  const saved = { log: console.log, warn: console.warn, error: console.error };
  console.log = console.warn = () => {};
  return Promise.resolve().then(fn).finally(() => Object.assign(console, saved));
}

// Builds a RayTracer of the reference for one case: {scene | preset, W, H, spp, depth, aa, tonemap, exposure, gamma, denoise,
// strength, perm (256 ints, optional)}
export function buildCase(ref, c) {
  const { RayTracer, Vec3 } = ref;
  const rt = new RayTracer(fakeCanvas(c.W, c.H));
  if (c.preset) rt.loadPreset(c.preset);
  else if (!rt.loadFromJSON(JSON.parse(JSON.stringify(c.scene)))) throw new Error('reference loadFromJSON returned false');
  rt.updateRenderSettings({ maxBounces: c.depth, samples: c.spp, gamma: c.gamma ?? 2.2, exposure: c.exposure ?? 1.0,
    toneMapping: c.tonemap ?? 'reinhard', antiAliasing: c.aa ?? 'supersampling', denoising: !!c.denoise, denoiseStrength: c.strength ?? 0.5 });
  const bg = c.scene?.background;
  if (bg && (bg.type === 'solid' || bg.type === 'hdri')) {
    // Deviation D1 (INTEGRATION.md §6): js/scene-loader.js:43,45 bind the background FACTORY instead of calling it, so the
    // loaded world returns a function where a colour is expected (NaN -> black) until the UI re-installs the background
    // (ui-controller.js:181).  The oracle and libbrt implement the intended behaviour; the harness installs it the way
    // ray-tracer.js:573-576 does, honouring the JSON colour.
    const col = bg.color ?? [0.1, 0.1, 0.1];
    rt.world.background = bg.type === 'solid' ? rt.world.solidBackground(new Vec3(col[0], col[1], col[2])) : rt.world.hdriBackground();
  }
  if (c.perm && rt.world.cloudNoise) {                              // world.cloudNoise.p is random per World (noise.js:7-17): an input here
    const p = rt.world.cloudNoise.p;
    for (let i = 0; i < 256; i++) { p[i] = c.perm[i]; p[256 + i] = c.perm[i]; }
  }
  return rt;
}

// Seeded render of one case through RayTracer.render(); returns {rgba, linear, float}
export async function renderSeeded(ref, c) {
  const rt = buildCase(ref, c);
  const W = c.W, H = c.H;
  const stream = new PhiloxStream(c.seed);
  const proto = Object.getPrototypeOf(rt);
  const origAA = proto.getAntiAliasSample, origTM = proto.toneMap, origGC = proto.gammaCorrect, origRandom = Math.random;
  const linear = new Float64Array(W * H * 3), fdat = new Float64Array(W * H * 3);
  let cur = 0;
  rt.getAntiAliasSample = function (i, j, s) { cur = (H - 1 - j) * W + i; stream.restart(cur, s); return origAA.call(this, i, j, s); };
  rt.toneMap = function (color) { linear.set([color.x, color.y, color.z], cur * 3); return origTM.call(this, color); };
  rt.gammaCorrect = function (color) { const r = origGC.call(this, color); fdat.set([r.x, r.y, r.z], cur * 3); return r; };
  Math.random = () => stream.next();
  globalThis.window.renderCancelled = false;
  try { await quiet(() => rt.render(() => {})); } finally { Math.random = origRandom; }
  return { rgba: Array.from(rt.imageData.data), linear: Array.from(linear), float: Array.from(fdat) };
}

// ------------------------------------------------------------------------------------------------ CLI
function args() {
  const a = process.argv.slice(2), o = {};
  for (let i = 0; i < a.length; i++) if (a[i].startsWith('--')) { const k = a[i].slice(2); o[k] = (i + 1 < a.length && !a[i + 1].startsWith('--')) ? a[++i] : true; }
  return o;
}

async function main() {
  const o = args();
  if (!o.ref) { console.error('usage: node baseline/run_ref.mjs --ref <dir with the reference js/> (--scene file.json | --preset name) --width W --height H [--samples n --bounces d] [--seed N --out file.json | --time-rows N | --as-shipped]'); process.exit(2); }
  const ref = await quiet(() => loadReference(o.ref));
  const c = { W: +(o.width ?? 600), H: +(o.height ?? 400), spp: +(o.samples ?? 4), depth: +(o.bounces ?? 5), aa: o.aa, tonemap: o.tonemap,
    exposure: o.exposure ? +o.exposure : undefined, gamma: o.gamma ? +o.gamma : undefined, denoise: !!o.denoise, strength: o.strength ? +o.strength : undefined };
  if (o.preset) c.preset = o.preset; else c.scene = JSON.parse(fs.readFileSync(o.scene, 'utf8'));
  if (o.seed !== undefined) {
    c.seed = o.seed;
    const out = await renderSeeded(ref, c);
    fs.writeFileSync(o.out ?? 'reference_render.json', JSON.stringify(out));
    return;
  }
  const rt = await quiet(() => buildCase(ref, c));
  if (rt.camera && !c.preset && !(c.scene.camera && c.scene.camera.resolution)) rt.resizeCanvas(c.W, c.H);   // aspect = W/H as the UI path does (:505)
  if (o['as-shipped']) {
    const t0 = performance.now();
    await quiet(() => rt.render(() => {}));
    const dt = (performance.now() - t0) / 1e3, n = c.W * c.H * (c.aa === 'none' ? 1 : c.spp);
    console.log(JSON.stringify({ mode: 'RayTracer.render() as shipped', msamples_per_s: n / dt / 1e6, seconds: dt, sample: `full ${c.W}x${c.H} frame at ${c.spp} spp = ${n} path samples, 1 thread, node ${process.version}` }));
    return;
  }
  const rows = Math.max(1, Math.min(c.H, +(o['time-rows'] ?? 8)));
  let n = 0;
  const t0 = performance.now();
  await quiet(() => {
    for (let k = 0; k < rows; k++) {
      const j = Math.min(c.H - 1, Math.floor((k + 0.5) * c.H / rows));                // rows spread evenly over the frame
      for (let i = 0; i < c.W; i++) for (let s = 0; s < c.spp; s++) {
        const sm = rt.getAntiAliasSample(i, j, s);
        rt.rayColor(rt.camera.getRay(sm.u, sm.v), rt.maxBounces);
        n++;
      }
    }
  });
  const dt = (performance.now() - t0) / 1e3;
  console.log(JSON.stringify({ mode: 'tight loop getAntiAliasSample -> getRay -> rayColor', msamples_per_s: n / dt / 1e6, seconds: dt,
    sample: `${rows} full-width rows spread evenly over ${c.W}x${c.H} at ${c.spp} spp = ${n} path samples, ${dt.toFixed(1)} s, 1 thread, node ${process.version}` }));
}

if (import.meta.url === pathToFileURL(process.argv[1] ?? '').href) main().catch((e) => { console.error(e); process.exit(1); });
