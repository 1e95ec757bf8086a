// mock_world.mjs — stand-ins for the reference's live objects, for machines without the reference checkout (the GPU box).
//
// napi/raytracer_gpu.mjs reads a World the reference built: class names through `constructor.name` and the members listed
// below (js/geometry.js, js/materials.js, js/lights.js, js/camera.js:14-35, js/world.js:9-18, js/ray-tracer.js:16-40).  These
// classes carry exactly those members and nothing else — no hit(), no scatter(): the renderer is libbrt.  buildRayTracer()
// fills them from a plain description (what brt_scene_get_flat reports for a scene the native loader ingested), so that the
// shim's image can be compared with the ctypes binding's image of the same scene.
class Vec3 { constructor(x, y, z) { this.x = x; this.y = y; this.z = z; } }
const v = (a) => new Vec3(a[0], a[1], a[2]);

class Lambertian { constructor(albedo) { this.albedo = albedo; } }
class Metal { constructor(albedo, roughness) { this.albedo = albedo; this.roughness = roughness; } }
class Dielectric { constructor(refractionIndex) { this.refractionIndex = refractionIndex; } }
class Emissive { constructor(color, intensity) { this.color = color; this.intensity = intensity; } }

class Sphere { constructor(center, radius, material) { this.center = center; this.radius = radius; this.material = material; } }
class Plane { constructor(point, normal, material) { this.point = point; this.normal = normal; this.material = material; } }
class Box { constructor(min, max, material) { this.min = min; this.max = max; this.material = material; } }
class Triangle { constructor(v0, v1, v2, material) { this.v0 = v0; this.v1 = v1; this.v2 = v2; this.material = material; } }
class TriangleMesh {                                   // no `material` member: the triangles hold it (geometry.js:231)
  constructor(tris, material) { this.triangles = tris.map((t) => new Triangle(v(t.slice(0, 3)), v(t.slice(3, 6)), v(t.slice(6, 9)), material)); }
}
class PointLight { constructor(position, color, intensity) { this.position = position; this.color = color; this.intensity = intensity; } }
class DirectionalLight { constructor(direction, color, intensity) { this.direction = direction; this.color = color; this.intensity = intensity; } }

class World {
  constructor() {
    this.objects = []; this.lights = []; this.skyIntensity = 1.0;
    this.cloudNoise = { p: Array.from({ length: 512 }, (_, i) => i % 256) };
    this.background = this.skyGradient.bind(this);
  }
  skyGradient() { return new Vec3(0, 0, 0); }
  proceduralSky() { return new Vec3(0, 0, 0); }
  solidBackground(color) { return () => color; }
  hdriBackground() { return () => new Vec3(0, 0, 0); }
}

class Camera {
  constructor(c) {
    for (const k of ['origin', 'lowerLeftCorner', 'horizontal', 'vertical', 'u', 'v', 'w']) this[k] = v(c[k]);
    this.lensRadius = c.lensRadius; this.type = c.type;
  }
}

function material(m) {
  if (m.kind === 'Metal') return new Metal(v(m.color), m.param);
  if (m.kind === 'Dielectric') return new Dielectric(m.param);
  if (m.kind === 'Emissive') return new Emissive(v(m.color), m.param);
  return new Lambertian(v(m.color));
}

class RayTracer {                                      // the members render() and the UI setters touch (ray-tracer.js:16-40, 554-585)
  constructor(canvas) {
    this.canvas = canvas; this.ctx = canvas.getContext('2d');
    this.width = canvas.width; this.height = canvas.height;
    this.imageData = this.ctx.createImageData(this.width, this.height);
    this.maxBounces = 5; this.samples = 4; this.gamma = 2.2; this.exposure = 1.0; this.toneMapping = 'reinhard';
    this.antiAliasing = 'supersampling'; this.denoising = false; this.denoiseStrength = 0.5;
    this.world = new World(); this.camera = null;
  }
  updateBackground(type, intensity = 1.0) {
    this.world.skyIntensity = intensity;
    if (type === 'solid') this.world.background = this.world.solidBackground(new Vec3(0.1, 0.1, 0.1));
    else if (type === 'hdri') this.world.background = this.world.hdriBackground();
    else if (type === 'procedural_sky') this.world.background = this.world.proceduralSky.bind(this.world);
    else this.world.background = this.world.skyGradient.bind(this.world);
  }
  loadPreset(name) { this.world = new World(); }
  loadFromJSON(json) { return false; }
  async render(onProgress) { throw new Error('the CPU loop is not part of the mock'); }
}

// desc: { objects: [{ kind, material: { kind, color, param }, a, b, c, tris }], lights: [{ kind, v, color, intensity }], camera, perm }
function buildRayTracer(canvas, desc) {
  const rt = new RayTracer(canvas);
  for (const o of desc.objects) {
    const m = material(o.material);
    if (o.kind === 'Sphere') rt.world.objects.push(new Sphere(v(o.a), o.b[0], m));
    else if (o.kind === 'Plane') rt.world.objects.push(new Plane(v(o.a), v(o.b), m));
    else if (o.kind === 'Box') rt.world.objects.push(new Box(v(o.a), v(o.b), m));
    else if (o.kind === 'Triangle') rt.world.objects.push(new Triangle(v(o.a), v(o.b), v(o.c), m));
    else rt.world.objects.push(new TriangleMesh(o.tris, m));
  }
  for (const l of desc.lights) rt.world.lights.push(l.kind === 'PointLight' ? new PointLight(v(l.v), v(l.color), l.intensity) : new DirectionalLight(v(l.v), v(l.color), l.intensity));
  rt.camera = new Camera(desc.camera);
  if (desc.perm) rt.world.cloudNoise.p = desc.perm.concat(desc.perm);
  return rt;
}

export { RayTracer, buildRayTracer };
