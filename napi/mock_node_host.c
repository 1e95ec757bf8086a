/* mock_node_host.c — a stand-in for Node.js that is just enough of an N-API host to LOAD and DRIVE brt_addon.node.
 *
 * TEST INFRASTRUCTURE.  This image has no Node.js, so the addon could otherwise only be compiled, never run.  This
 * program implements the 36 napi_* functions the addon imports over a tiny tagged-value heap, dlopen()s the addon
 * (the napi_* symbols resolve against this executable, as they do against `node`), calls napi_register_module_v1 and
 * then drives the exported functions the way napi/raytracer_gpu.mjs does:
 *     create -> loadSceneJSON -> setRenderParams -> render(ctx, Uint8ClampedArray, onProgress) -> Promise
 * and writes the pixels the addon produced to a file, so a test can compare them byte for byte with the Python binding's
 * render of the same scene, seed and settings (tests/test_napi_mock.py).
 * Differences from Node, by design: async work runs synchronously inside napi_queue_async_work and thread-safe function
 * calls are delivered immediately on the calling thread; values are never garbage-collected.
 *
 *   cc -O1 -rdynamic -Inapi -Iinclude napi/mock_node_host.c -ldl -o napi/mock_node_host
 *   napi/mock_node_host napi/brt_addon.node scene.json W H spp depth seed out.rgba
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "node_api_min.h"

typedef struct Prop { char* name; struct Val* v; struct Prop* next; } Prop;
typedef struct Val {
    napi_valuetype type;
    double num; bool b; char* str; void* ext;
    Prop* props;
    napi_callback fn; void* fn_data;                               /* function */
    int is_typed; napi_typedarray_type ta_type; void* ta_data; size_t ta_len;   /* typed array (type == napi_object) */
    int is_error; int is_promise; int settled; struct Val* result;              /* promise: settled 1 = resolved, 2 = rejected */
    int is_array; struct Val** elems; uint32_t n_elems;                         /* Array (type == napi_object) */
    struct Val* (*native)(struct Val** argv, size_t argc);         /* host-side JS function (onProgress) */
} Val;
struct napi_callback_info__ { size_t argc; Val** argv; void* data; };
struct napi_deferred__ { Val* promise; };
struct napi_async_work__ { napi_async_execute_callback exec; napi_async_complete_callback done; void* data; };
struct napi_threadsafe_function__ { Val* fn; napi_threadsafe_function_call_js call_js; void* ctx; };
struct napi_env__ { Val* pending; };

static Val* mk(napi_valuetype t) { Val* v = (Val*)calloc(1, sizeof(Val)); v->type = t; return v; }
static Val UNDEF = { napi_undefined };
#define V(x) ((Val*)(x))
#define NV(x) ((napi_value)(x))

napi_status napi_get_undefined(napi_env e, napi_value* r) { (void)e; *r = NV(&UNDEF); return napi_ok; }
napi_status napi_create_object(napi_env e, napi_value* r) { (void)e; *r = NV(mk(napi_object)); return napi_ok; }
napi_status napi_create_int32(napi_env e, int32_t v, napi_value* r) { (void)e; Val* x = mk(napi_number); x->num = v; *r = NV(x); return napi_ok; }
napi_status napi_create_double(napi_env e, double v, napi_value* r) { (void)e; Val* x = mk(napi_number); x->num = v; *r = NV(x); return napi_ok; }
napi_status napi_get_boolean(napi_env e, bool v, napi_value* r) { (void)e; Val* x = mk(napi_boolean); x->b = v; *r = NV(x); return napi_ok; }
napi_status napi_create_string_utf8(napi_env e, const char* s, size_t n, napi_value* r) {
    (void)e; Val* x = mk(napi_string); if (n == NAPI_AUTO_LENGTH) n = strlen(s);
    x->str = (char*)malloc(n + 1); memcpy(x->str, s, n); x->str[n] = 0; *r = NV(x); return napi_ok;
}
napi_status napi_typeof(napi_env e, napi_value v, napi_valuetype* r) { (void)e; *r = V(v)->type; return napi_ok; }
napi_status napi_is_array(napi_env e, napi_value v, bool* r) { (void)e; *r = V(v)->is_array != 0; return napi_ok; }
napi_status napi_get_array_length(napi_env e, napi_value v, uint32_t* r) { (void)e; if (!V(v)->is_array) return napi_array_expected; *r = V(v)->n_elems; return napi_ok; }
napi_status napi_get_element(napi_env e, napi_value v, uint32_t i, napi_value* r) { (void)e; *r = NV(V(v)->is_array && i < V(v)->n_elems ? V(v)->elems[i] : &UNDEF); return napi_ok; }
napi_status napi_get_value_int32(napi_env e, napi_value v, int32_t* r) { (void)e; if (V(v)->type != napi_number) return napi_number_expected; *r = (int32_t)V(v)->num; return napi_ok; }
napi_status napi_get_value_double(napi_env e, napi_value v, double* r) { (void)e; if (V(v)->type != napi_number) return napi_number_expected; *r = V(v)->num; return napi_ok; }
napi_status napi_get_value_bool(napi_env e, napi_value v, bool* r) { (void)e; if (V(v)->type != napi_boolean) return napi_boolean_expected; *r = V(v)->b; return napi_ok; }
napi_status napi_get_value_string_utf8(napi_env e, napi_value v, char* buf, size_t bufsize, size_t* r) {
    (void)e; if (V(v)->type != napi_string) return napi_string_expected;
    size_t n = strlen(V(v)->str);
    if (!buf) { if (r) *r = n; return napi_ok; }
    size_t c = n < bufsize - 1 ? n : bufsize - 1; memcpy(buf, V(v)->str, c); buf[c] = 0; if (r) *r = c; return napi_ok;
}
napi_status napi_set_named_property(napi_env e, napi_value o, const char* name, napi_value v) {
    (void)e; Prop* p = (Prop*)calloc(1, sizeof(Prop)); p->name = strdup(name); p->v = V(v); p->next = V(o)->props; V(o)->props = p; return napi_ok;
}
static Val* find(Val* o, const char* name) { for (Prop* p = o->props; p; p = p->next) if (!strcmp(p->name, name)) return p->v; return NULL; }
napi_status napi_get_named_property(napi_env e, napi_value o, const char* name, napi_value* r) { (void)e; Val* v = find(V(o), name); *r = NV(v ? v : &UNDEF); return napi_ok; }
napi_status napi_has_named_property(napi_env e, napi_value o, const char* name, bool* r) { (void)e; *r = find(V(o), name) != NULL; return napi_ok; }
napi_status napi_create_external(napi_env e, void* data, napi_finalize fin, void* hint, napi_value* r) { (void)e; (void)fin; (void)hint; Val* x = mk(napi_external); x->ext = data; *r = NV(x); return napi_ok; }
napi_status napi_get_value_external(napi_env e, napi_value v, void** r) { (void)e; if (V(v)->type != napi_external) return napi_invalid_arg; *r = V(v)->ext; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env e, napi_value v, napi_typedarray_type* t, size_t* len, void** data, napi_value* ab, size_t* off) {
    (void)e; if (!V(v)->is_typed) return napi_invalid_arg;
    if (t) *t = V(v)->ta_type; if (len) *len = V(v)->ta_len; if (data) *data = V(v)->ta_data; if (ab) *ab = NULL; if (off) *off = 0; return napi_ok;
}
napi_status napi_create_function(napi_env e, const char* name, size_t n, napi_callback cb, void* data, napi_value* r) {
    (void)e; (void)name; (void)n; Val* x = mk(napi_function); x->fn = cb; x->fn_data = data; *r = NV(x); return napi_ok;
}
napi_status napi_get_cb_info(napi_env e, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
    (void)e;
    if (argv && argc) { for (size_t i = 0; i < *argc; i++) argv[i] = NV(i < info->argc ? info->argv[i] : &UNDEF); }
    if (argc) *argc = info->argc;
    if (this_arg) *this_arg = NV(&UNDEF);
    if (data) *data = info->data;
    return napi_ok;
}
napi_status napi_create_error(napi_env e, napi_value code, napi_value msg, napi_value* r) {
    (void)e; Val* x = mk(napi_object); x->is_error = 1;
    napi_set_named_property(e, NV(x), "message", msg); if (code) napi_set_named_property(e, NV(x), "code", code); *r = NV(x); return napi_ok;
}
napi_status napi_throw_error(napi_env e, const char* code, const char* msg) {
    napi_value c, m, err; napi_create_string_utf8(e, code ? code : "", NAPI_AUTO_LENGTH, &c); napi_create_string_utf8(e, msg ? msg : "", NAPI_AUTO_LENGTH, &m);
    napi_create_error(e, c, m, &err); e->pending = V(err); return napi_ok;
}
napi_status napi_create_promise(napi_env e, napi_deferred* d, napi_value* p) {
    (void)e; Val* x = mk(napi_object); x->is_promise = 1; *d = (napi_deferred)calloc(1, sizeof(**d)); (*d)->promise = x; *p = NV(x); return napi_ok;
}
napi_status napi_resolve_deferred(napi_env e, napi_deferred d, napi_value v) { (void)e; d->promise->settled = 1; d->promise->result = V(v); free(d); return napi_ok; }
napi_status napi_reject_deferred(napi_env e, napi_deferred d, napi_value v) { (void)e; d->promise->settled = 2; d->promise->result = V(v); free(d); return napi_ok; }
napi_status napi_create_reference(napi_env e, napi_value v, uint32_t n, napi_ref* r) { (void)e; (void)n; *r = (napi_ref)v; return napi_ok; }
napi_status napi_delete_reference(napi_env e, napi_ref r) { (void)e; (void)r; return napi_ok; }
napi_status napi_call_function(napi_env e, napi_value recv, napi_value fn, size_t argc, const napi_value* argv, napi_value* r) {
    (void)recv; Val* f = V(fn); if (f->type != napi_function) return napi_function_expected;
    Val* out = &UNDEF;
    if (f->native) out = f->native((Val**)argv, argc);
    else { struct napi_callback_info__ info = { argc, (Val**)argv, f->fn_data }; napi_value x = f->fn(e, &info); out = x ? V(x) : &UNDEF; }
    if (r) *r = NV(out);
    return napi_ok;
}
napi_status napi_create_async_work(napi_env e, napi_value res, napi_value name, napi_async_execute_callback ex, napi_async_complete_callback done, void* data, napi_async_work* r) {
    (void)e; (void)res; (void)name; *r = (napi_async_work)calloc(1, sizeof(**r)); (*r)->exec = ex; (*r)->done = done; (*r)->data = data; return napi_ok;
}
napi_status napi_queue_async_work(napi_env e, napi_async_work w) { w->exec(e, w->data); w->done(e, napi_ok, w->data); return napi_ok; }   /* synchronous */
napi_status napi_delete_async_work(napi_env e, napi_async_work w) { (void)e; (void)w; return napi_ok; }   /* the job struct is still on the caller's stack frame */
napi_status napi_create_threadsafe_function(napi_env e, napi_value fn, napi_value res, napi_value name, size_t q, size_t th, void* fd, napi_finalize fcb,
                                            void* ctx, napi_threadsafe_function_call_js cjs, napi_threadsafe_function* r) {
    (void)e; (void)res; (void)name; (void)q; (void)th; (void)fd; (void)fcb;
    *r = (napi_threadsafe_function)calloc(1, sizeof(**r)); (*r)->fn = V(fn); (*r)->call_js = cjs; (*r)->ctx = ctx; return napi_ok;
}
static struct napi_env__ ENV;
napi_status napi_call_threadsafe_function(napi_threadsafe_function f, void* data, napi_threadsafe_function_call_mode m) { (void)m; f->call_js(&ENV, NV(f->fn), f->ctx, data); return napi_ok; }
napi_status napi_release_threadsafe_function(napi_threadsafe_function f, napi_threadsafe_function_release_mode m) { (void)m; free(f); return napi_ok; }

/* ------------------------------------------------------------------------------------------------ the "JavaScript" side */
static int n_progress = 0; static double last_progress = -1;
static Val* on_progress(Val** argv, size_t argc) { if (argc >= 1) { n_progress++; last_progress = argv[0]->num; } return &UNDEF; }
static Val* num(double x) { Val* v = mk(napi_number); v->num = x; return v; }
static Val* call(Val* exports, const char* name, Val** argv, size_t argc) {
    Val* f = find(exports, name);
    if (!f) { fprintf(stderr, "addon does not export %s\n", name); exit(3); }
    ENV.pending = NULL;
    napi_value r; napi_call_function(&ENV, NV(&UNDEF), NV(f), argc, (const napi_value*)argv, &r);
    if (ENV.pending) { Val* m = find(ENV.pending, "message"); Val* c = find(ENV.pending, "code"); fprintf(stderr, "%s threw %s: %s\n", name, c ? c->str : "?", m ? m->str : "?"); exit(4); }
    return V(r);
}

int main(int argc, char** argv) {
    if (argc < 9) { fprintf(stderr, "usage: %s addon.node scene.json W H spp depth seed out.rgba\n", argv[0]); return 2; }
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    napi_value (*reg)(napi_env, napi_value) = (napi_value (*)(napi_env, napi_value))dlsym(h, "napi_register_module_v1");
    if (!reg) { fprintf(stderr, "napi_register_module_v1 missing\n"); return 2; }
    Val* exports = mk(napi_object);
    reg(&ENV, NV(exports));
    FILE* f = fopen(argv[2], "rb"); if (!f) { perror(argv[2]); return 2; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    char* text = (char*)malloc(n + 1); if (fread(text, 1, n, f) != (size_t)n) return 2; text[n] = 0; fclose(f);
    int W = atoi(argv[3]), H = atoi(argv[4]), spp = atoi(argv[5]), depth = atoi(argv[6]); double seed = atof(argv[7]);

    /* create(0), or with BRT_MOCK_DEVICES="0,1,..." create([0, 1, ...]): one ctx spanning several GPUs */
    Val* a1[1] = { num(0) };
    const char* devs = getenv("BRT_MOCK_DEVICES");
    if (devs && *devs) {
        Val* arr = mk(napi_object); arr->is_array = 1; arr->elems = (Val**)calloc(16, sizeof(Val*));
        char* copy = strdup(devs);
        for (char* tok = strtok(copy, ","); tok && arr->n_elems < 16; tok = strtok(NULL, ",")) arr->elems[arr->n_elems++] = num(atof(tok));
        a1[0] = arr;
    }
    Val* ctx = call(exports, "create", a1, 1);
    Val* js = mk(napi_string); js->str = text;
    Val* a4[4] = { ctx, js, num(W), num(H) };
    Val* info = call(exports, "loadSceneJSON", a4, 4);
    printf("loadSceneJSON: hasCamera=%d\n", (int)find(info, "hasCamera")->b);
    Val* p = mk(napi_object);
    napi_set_named_property(&ENV, NV(p), "width", NV(num(W))); napi_set_named_property(&ENV, NV(p), "height", NV(num(H)));
    napi_set_named_property(&ENV, NV(p), "samples", NV(num(spp))); napi_set_named_property(&ENV, NV(p), "maxBounces", NV(num(depth)));
    napi_set_named_property(&ENV, NV(p), "seed", NV(num(seed))); napi_set_named_property(&ENV, NV(p), "sppBatch", NV(num(spp > 4 ? spp / 4 : 1)));
    Val* a2[2] = { ctx, p };
    call(exports, "setRenderParams", a2, 2);
    Val* px = mk(napi_object); px->is_typed = 1; px->ta_type = napi_uint8_clamped_array; px->ta_len = (size_t)W * H * 4; px->ta_data = calloc(px->ta_len, 1);
    Val* cb = mk(napi_function); cb->native = on_progress;
    Val* a3[3] = { ctx, px, cb };
    Val* promise = call(exports, "render", a3, 3);
    if (!promise->is_promise || promise->settled != 1) {
        Val* m = promise->result ? find(promise->result, "message") : NULL;
        fprintf(stderr, "render promise not resolved (state %d): %s\n", promise->settled, m ? m->str : "?"); return 5;
    }
    Val* a5[1] = { ctx };
    Val* st = call(exports, "stats", a5, 1);
    printf("render resolved; onProgress calls=%d last=%.3f; stats.samples=%.0f launches=%.0f devices=%.0f\n", n_progress, last_progress,
           find(st, "samples")->num, find(st, "launches")->num, find(st, "devices")->num);
    FILE* o = fopen(argv[8], "wb"); if (!o) { perror(argv[8]); return 2; }
    fwrite(px->ta_data, 1, px->ta_len, o); fclose(o);
    return 0;
}
