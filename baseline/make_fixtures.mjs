// make_fixtures.mjs — turns the reference's own output into golden vectors for the oracle.
//
//   node baseline/make_fixtures.mjs --stage /path/to/BlenderRayTracer     # once: copies js/ to baseline/_ref/js (+ {"type":"module"})
//   node baseline/make_fixtures.mjs                                        # writes tests/golden/reference_vectors.json
//   python -m pytest tests/test_reference_pin.py                           # the oracle against the reference's own numbers
//
// Cases = tests/golden/reference_cases.json (written by tests/golden/make_reference_cases.py: the 13 cases of the
// second-port cross-check — both fixtures, the four presets, the four backgrounds, all AA / tone-map modes, denoise, the
// orthographic camera — with explicit seeds and Perlin tables).  Each is rendered by the UNMODIFIED reference through
// baseline/run_ref.mjs with Math.random replaced by the oracle's Philox stream.
import fs from 'node:fs';
import path from 'node:path';
import { fileURLToPath } from 'node:url';
import { loadReference, renderSeeded } from './run_ref.mjs';

const HERE = path.dirname(fileURLToPath(import.meta.url));
const ROOT = path.resolve(HERE, '..');
const REF = path.join(HERE, '_ref');

const a = process.argv.slice(2);
if (a[0] === '--stage') {
  const src = path.join(a[1], 'js');
  fs.mkdirSync(path.join(REF, 'js'), { recursive: true });
  for (const f of fs.readdirSync(src)) if (f.endsWith('.js')) fs.copyFileSync(path.join(src, f), path.join(REF, 'js', f));
  fs.writeFileSync(path.join(REF, 'package.json'), '{"type":"module"}\n');      // the reference's .js files are ES modules
  console.log(`staged ${src} -> ${path.join(REF, 'js')} (git-ignored; nothing of the reference is committed)`);
  process.exit(0);
}

const saved = { log: console.log, warn: console.warn };
console.log = console.warn = () => {};
const ref = await loadReference(path.join(REF, 'js'));
Object.assign(console, saved);
const cases = [];
for (const fn of ['reference_cases.json', 'reference_cases_extra.json']) {       // the 13 second-port cases + the 7 BASELINE-shaped ones
    const p = path.join(ROOT, 'tests', 'golden', fn);
    if (fs.existsSync(p)) cases.push(...JSON.parse(fs.readFileSync(p, 'utf8')));
}
const out = { generator: `baseline/make_fixtures.mjs, node ${process.version}`, cases: {} };
for (const c of cases) {
  const r = await renderSeeded(ref, c);
  out.cases[c.name] = { W: c.W, H: c.H, ...r };
  console.log(`${c.name}: ${c.W}x${c.H}, mean linear ${(r.linear.reduce((s, v) => s + v, 0) / r.linear.length).toFixed(6)}`);
}
fs.writeFileSync(path.join(ROOT, 'tests', 'golden', 'reference_vectors.json'), JSON.stringify(out));
console.log('wrote tests/golden/reference_vectors.json');
