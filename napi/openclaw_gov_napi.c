/*
 * napi/openclaw_gov_napi.c -- thin N-API (node_api.h, C ABI, ABI-stable) shim over libopenclaw_gov.so.
 *
 * This is the binding the TypeScript plugin (@vainplex/openclaw-governance, src/hooks.ts,
 * src/redaction/registry.ts) would call instead of V8 RegExp / node:crypto on the hot path.
 * It is NOT built in this repository's image (no Node.js, no node_api.h here or on the GPU box);
 * it is kept compilable against any Node >= 22 header set:
 *
 *   cc -shared -fPIC -I"$(node -p 'process.config.variables.node_prefix')/include/node" \
 *      -Iinclude napi/openclaw_gov_napi.c -Lvainplex_openclaw_b200 -lopenclaw_gov \
 *      -Wl,-rpath,'$ORIGIN' -o openclaw_gov.node
 *
 * JS surface (all synchronous -- the sync hooks before_message_write / tool_result_persist cannot
 * await, src/hooks.ts:360-365; the async bulk path wraps scanBatch in a libuv worker on the JS side):
 *   init(device?: number): void
 *   createRuleset(rules: {source: string, flags?: number, category?: number}[]): {handle: External, status: Int32Array}
 *   scanBatch(handle, bytes: Uint8Array, offsets: Uint32Array): {words: BigUint64Array, hits: Uint32Array}  // hits: msg,rule pairs
 *   findMatchesBatch(handle, bytes, offsets): Uint32Array  // 6 words per span: msg, rule, startByte, endByte, start16, end16
 *   redactBatch(handle, bytes, offsets): {bytes, offsets, spans, digests}  // scanString for a batch, spliced on the device
 *   sha256Batch(bytes: Uint8Array, offsets: BigUint64Array): Uint8Array  // 32 bytes per item
 *   merkleRoot(bytes: Uint8Array, offsets: BigUint64Array): Uint8Array   // 32 bytes
 * Every failure throws a JS Error carrying cg_last_error(); the reference's try/catch + failMode
 * blocks (src/hooks.ts:232-241, src/redaction/hooks.ts:193-204) decide what happens next.
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>

#include "openclaw_gov.h"

#define NAPI_CALL(env, call) do { if ((call) != napi_ok) { napi_throw_error((env), NULL, "N-API call failed: " #call); return NULL; } } while (0)
#define CG_CALL(env, call) do { int _rc = (call); if (_rc != CG_OK) { napi_throw_error((env), NULL, cg_last_error()); return NULL; } } while (0)

static void ruleset_finalize(napi_env env, void *data, void *hint) { (void)env; (void)hint; cg_ruleset_destroy((cg_ruleset *)data); }

static napi_value js_init(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1]; int32_t dev = -1;
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc >= 1) napi_get_value_int32(env, argv[0], &dev);
  CG_CALL(env, cg_init(dev));
  return NULL;
}

static napi_value js_create_ruleset(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value argv[1]; uint32_t n = 0;
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_CALL(env, napi_get_array_length(env, argv[0], &n));
  cg_rule *rules = (cg_rule *)calloc(n ? n : 1, sizeof(cg_rule));
  char **bufs = (char **)calloc(n ? n : 1, sizeof(char *));
  for (uint32_t i = 0; i < n; i++) {
    napi_value r, v; size_t len = 0; bool has;
    napi_get_element(env, argv[0], i, &r);
    napi_get_named_property(env, r, "source", &v);
    napi_get_value_string_utf8(env, v, NULL, 0, &len);
    bufs[i] = (char *)malloc(len + 1);
    napi_get_value_string_utf8(env, v, bufs[i], len + 1, &len);       /* lone surrogates -> U+FFFD, like Buffer.from */
    rules[i].source = bufs[i]; rules[i].source_len = (uint32_t)len;
    if (napi_has_named_property(env, r, "flags", &has) == napi_ok && has) { napi_get_named_property(env, r, "flags", &v); napi_get_value_uint32(env, v, &rules[i].flags); }
    rules[i].category = CG_CAT_CUSTOM;
    if (napi_has_named_property(env, r, "category", &has) == napi_ok && has) { napi_get_named_property(env, r, "category", &v); napi_get_value_uint32(env, v, &rules[i].category); }
  }
  napi_value status_ab, status; void *status_data;
  napi_create_arraybuffer(env, (size_t)n * 4, &status_data, &status_ab);
  napi_create_typedarray(env, napi_int32_array, n, status_ab, 0, &status);
  cg_ruleset *rs = NULL;
  int rc = cg_ruleset_create(rules, n, CG_OPT_STRIDE_AUTO, &rs, (int32_t *)status_data);
  for (uint32_t i = 0; i < n; i++) free(bufs[i]);
  free(bufs); free(rules);
  if (rc != CG_OK) { napi_throw_error(env, NULL, cg_last_error()); return NULL; }
  napi_value out, handle;
  napi_create_object(env, &out);
  napi_create_external(env, rs, ruleset_finalize, NULL, &handle);
  napi_set_named_property(env, out, "handle", handle);
  napi_set_named_property(env, out, "status", status);
  return out;
}

static int get_u8(napi_env env, napi_value v, uint8_t **p, size_t *len) {
  napi_typedarray_type t; napi_value ab; size_t off;
  return napi_get_typedarray_info(env, v, &t, len, (void **)p, &ab, &off) == napi_ok;
}

static napi_value js_scan_batch(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3]; cg_ruleset *rs; uint8_t *bytes, *offp; size_t nb, no;
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_CALL(env, napi_get_value_external(env, argv[0], (void **)&rs));
  if (!get_u8(env, argv[1], &bytes, &nb) || !get_u8(env, argv[2], &offp, &no) || no == 0) { napi_throw_type_error(env, NULL, "scanBatch(handle, Uint8Array, Uint32Array)"); return NULL; }
  uint32_t n = (uint32_t)(no - 1);
  napi_value words_ab, words, hits_ab, hits, out; void *wdata, *hdata;
  napi_create_arraybuffer(env, (size_t)n * 8, &wdata, &words_ab);
  napi_create_typedarray(env, napi_biguint64_array, n, words_ab, 0, &words);
  uint32_t nh = 0, cap = 1024; cg_hit *tmp = (cg_hit *)malloc(sizeof(cg_hit) * cap);
  int rc = cg_scan_batch(rs, bytes, (const uint32_t *)offp, n, (uint64_t *)wdata, tmp, cap, &nh);
  if (rc == CG_ERR_CAPACITY) { cap = nh; tmp = (cg_hit *)realloc(tmp, sizeof(cg_hit) * cap); rc = cg_scan_batch(rs, bytes, (const uint32_t *)offp, n, (uint64_t *)wdata, tmp, cap, &nh); }
  if (rc != CG_OK) { free(tmp); napi_throw_error(env, NULL, cg_last_error()); return NULL; }
  napi_create_arraybuffer(env, (size_t)nh * 8, &hdata, &hits_ab);
  memcpy(hdata, tmp, (size_t)nh * 8); free(tmp);
  napi_create_typedarray(env, napi_uint32_array, (size_t)nh * 2, hits_ab, 0, &hits);
  napi_create_object(env, &out);
  napi_set_named_property(env, out, "words", words);
  napi_set_named_property(env, out, "hits", hits);
  return out;
}

static napi_value js_find_matches_batch(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3]; cg_ruleset *rs; uint8_t *bytes, *offp; size_t nb, no;
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_CALL(env, napi_get_value_external(env, argv[0], (void **)&rs));
  if (!get_u8(env, argv[1], &bytes, &nb) || !get_u8(env, argv[2], &offp, &no) || no == 0) { napi_throw_type_error(env, NULL, "findMatchesBatch(handle, Uint8Array, Uint32Array)"); return NULL; }
  uint32_t n = (uint32_t)(no - 1), ns = 0, cap = 256; cg_span *tmp = (cg_span *)malloc(sizeof(cg_span) * cap);
  int rc = cg_find_matches_batch(rs, bytes, (const uint32_t *)offp, n, tmp, cap, &ns);
  if (rc == CG_ERR_CAPACITY) { cap = ns; tmp = (cg_span *)realloc(tmp, sizeof(cg_span) * cap); rc = cg_find_matches_batch(rs, bytes, (const uint32_t *)offp, n, tmp, cap, &ns); }
  if (rc != CG_OK) { free(tmp); napi_throw_error(env, NULL, cg_last_error()); return NULL; }
  napi_value ab, arr; void *data;
  napi_create_arraybuffer(env, (size_t)ns * sizeof(cg_span), &data, &ab);
  memcpy(data, tmp, (size_t)ns * sizeof(cg_span)); free(tmp);
  napi_create_typedarray(env, napi_uint32_array, (size_t)ns * 6, ab, 0, &arr);
  return arr;
}

static napi_value js_sha256_batch(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2]; uint8_t *bytes, *offp; size_t nb, no;
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (!get_u8(env, argv[0], &bytes, &nb) || !get_u8(env, argv[1], &offp, &no) || no == 0) { napi_throw_type_error(env, NULL, "sha256Batch(Uint8Array, BigUint64Array)"); return NULL; }
  uint32_t n = (uint32_t)(no - 1);
  napi_value ab, arr; void *data;
  napi_create_arraybuffer(env, (size_t)n * 32, &data, &ab);
  CG_CALL(env, cg_sha256_batch(bytes, (const uint64_t *)offp, n, (uint8_t *)data));
  napi_create_typedarray(env, napi_uint8_array, (size_t)n * 32, ab, 0, &arr);
  return arr;
}

static napi_value js_merkle_root(napi_env env, napi_callback_info info) {
  size_t argc = 2; napi_value argv[2]; uint8_t *bytes, *offp; size_t nb, no;
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (!get_u8(env, argv[0], &bytes, &nb) || !get_u8(env, argv[1], &offp, &no) || no == 0) { napi_throw_type_error(env, NULL, "merkleRoot(Uint8Array, BigUint64Array)"); return NULL; }
  napi_value ab, arr; void *data;
  napi_create_arraybuffer(env, 32, &data, &ab);
  CG_CALL(env, cg_merkle_root(bytes, (const uint64_t *)offp, (uint64_t)(no - 1), (uint8_t *)data));
  napi_create_typedarray(env, napi_uint8_array, 32, ab, 0, &arr);
  return arr;
}

/* redactBatch(handle, bytes, offsets) -> { bytes: Uint8Array, offsets: Uint32Array, spans: Uint32Array (6 per span), digests: Uint8Array (32 per span) }
 * = RedactionEngine.scanString for a batch (engine.ts:74-85,165-181) with the vault's default placeholder; the TS side
 * stores the originals in the vault from spans + digests (vault.ts:75-128). */
static napi_value js_redact_batch(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value argv[3]; cg_ruleset *rs; uint8_t *bytes, *offp; size_t nb, no;
  NAPI_CALL(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_CALL(env, napi_get_value_external(env, argv[0], (void **)&rs));
  if (!get_u8(env, argv[1], &bytes, &nb) || !get_u8(env, argv[2], &offp, &no) || no == 0) { napi_throw_type_error(env, NULL, "redactBatch(handle, Uint8Array, Uint32Array)"); return NULL; }
  uint32_t n = (uint32_t)(no - 1), ns = 0, cap_spans = 256; uint64_t need = 0, cap_bytes = nb + 4096;
  napi_value off_ab, off_arr; void *off_data;
  napi_create_arraybuffer(env, (size_t)(n + 1) * 4, &off_data, &off_ab);
  uint8_t *out = NULL, *dig = NULL; cg_span *spans = NULL; int rc;
  for (;;) {
    out = (uint8_t *)realloc(out, cap_bytes + 64); spans = (cg_span *)realloc(spans, sizeof(cg_span) * cap_spans); dig = (uint8_t *)realloc(dig, (size_t)cap_spans * 32);
    rc = cg_redact_batch(rs, bytes, (const uint32_t *)offp, n, out, cap_bytes, &need, (uint32_t *)off_data, spans, cap_spans, &ns, dig);
    if (rc == CG_ERR_CAPACITY && (need > cap_bytes || ns > cap_spans)) { if (need > cap_bytes) cap_bytes = need; if (ns > cap_spans) cap_spans = ns; continue; }
    break;
  }
  if (rc != CG_OK) { free(out); free(spans); free(dig); napi_throw_error(env, NULL, cg_last_error()); return NULL; }
  napi_value res, b_ab, b_arr, s_ab, s_arr, d_ab, d_arr; void *p;
  napi_create_arraybuffer(env, (size_t)need, &p, &b_ab); memcpy(p, out, (size_t)need); napi_create_typedarray(env, napi_uint8_array, (size_t)need, b_ab, 0, &b_arr);
  napi_create_arraybuffer(env, (size_t)ns * sizeof(cg_span), &p, &s_ab); memcpy(p, spans, (size_t)ns * sizeof(cg_span)); napi_create_typedarray(env, napi_uint32_array, (size_t)ns * 6, s_ab, 0, &s_arr);
  napi_create_arraybuffer(env, (size_t)ns * 32, &p, &d_ab); memcpy(p, dig, (size_t)ns * 32); napi_create_typedarray(env, napi_uint8_array, (size_t)ns * 32, d_ab, 0, &d_arr);
  napi_create_typedarray(env, napi_uint32_array, (size_t)n + 1, off_ab, 0, &off_arr);
  free(out); free(spans); free(dig);
  napi_create_object(env, &res);
  napi_set_named_property(env, res, "bytes", b_arr); napi_set_named_property(env, res, "offsets", off_arr);
  napi_set_named_property(env, res, "spans", s_arr); napi_set_named_property(env, res, "digests", d_arr);
  return res;
}

static napi_value module_init(napi_env env, napi_value exports) {
  napi_property_descriptor d[] = {
    {"init", NULL, js_init, NULL, NULL, NULL, napi_default, NULL},
    {"createRuleset", NULL, js_create_ruleset, NULL, NULL, NULL, napi_default, NULL},
    {"scanBatch", NULL, js_scan_batch, NULL, NULL, NULL, napi_default, NULL},
    {"findMatchesBatch", NULL, js_find_matches_batch, NULL, NULL, NULL, napi_default, NULL},
    {"redactBatch", NULL, js_redact_batch, NULL, NULL, NULL, napi_default, NULL},
    {"sha256Batch", NULL, js_sha256_batch, NULL, NULL, NULL, napi_default, NULL},
    {"merkleRoot", NULL, js_merkle_root, NULL, NULL, NULL, napi_default, NULL},
  };
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, module_init)
