/**
 * napi/gpu-backend.ts -- the reference-side glue a maintainer of @vainplex/openclaw-governance would add
 * (packages/openclaw-governance/src/gpu-backend.ts).  It keeps the plugin's classes and call sites and swaps the two hot
 * loops and the hash for calls into the N-API addon built from napi/openclaw_gov_napi.c:
 *
 *   PatternRegistry.findMatches   src/redaction/registry.ts:212-242,288-316   -> findMatchesBatch (spans already resolved)
 *   matchesAny                    src/conditions/context.ts:9-25              -> scanBatch over every pattern of the policy set
 *   ResponseGate mustMatch / mustNotMatch  src/response-gate.ts:104-149       -> scanBatch
 *   tool param `matches`          src/conditions/tool.ts:36-45                -> scanBatch
 *   sha256                        src/util.ts:77-79, src/redaction/vault.ts:26-28 -> sha256Batch
 *   audit JSONL -> Merkle log     src/audit-trail.ts:151-179 (flush)          -> logAppendJsonl / logRoot / logFrontier
 *
 * Not compiled or run in this repository (no Node.js in the image or on the GPU box): it documents the binding;
 * vainplex_openclaw_b200/plugin.py is the executable mirror of the same wiring and is what the tests drive.
 */
import { createRequire } from "node:module";

const require = createRequire(import.meta.url);

type Addon = {
  init(device?: number): void;
  shutdown(): void;
  stats(): Record<string, number>;
  ruleCheck(source: string, flags?: number): number;
  createRuleset(rules: { source: string; flags?: number; category?: number }[]): { handle: unknown; status: Int32Array };
  scanBatch(handle: unknown, bytes: Uint8Array, offsets: Uint32Array): { words: BigUint64Array; hits: Uint32Array };
  scanOne(handle: unknown, bytes: Uint8Array): { word: bigint; rules: Uint32Array };
  findMatchesBatch(handle: unknown, bytes: Uint8Array, offsets: Uint32Array): Uint32Array;
  redactBatch(handle: unknown, bytes: Uint8Array, offsets: Uint32Array): { bytes: Uint8Array; offsets: Uint32Array; spans: Uint32Array; digests: Uint8Array };
  setPolicy(handle: unknown, rulePolicy: Uint32Array, ruleAction: Uint8Array): void;
  verdictBatch(handle: unknown, bytes: Uint8Array, offsets: Uint32Array): Uint32Array;
  sha256Batch(bytes: Uint8Array, offsets: BigUint64Array): Uint8Array;
  merkleRoot(bytes: Uint8Array, offsets: BigUint64Array): Uint8Array;
  logCreate(keepLeafDigests: boolean): unknown;
  logRestore(size: bigint, frontier: Uint8Array): unknown;
  logAppendJsonl(log: unknown, bytes: Uint8Array): bigint;
  logRoot(log: unknown): Uint8Array;
  logSize(log: unknown): bigint;
  logFrontier(log: unknown): Uint8Array;
  logProof(log: unknown, index: bigint): Uint8Array;
  logConsistency(log: unknown, firstSize: bigint): Uint8Array;
};

const addon: Addon = require("./openclaw_gov.node");
const FLAG_ICASE = 1;
const CATEGORY: Record<string, number> = { credential: 0, pii: 1, financial: 2, custom: 3 };
const enc = new TextEncoder();

/** Buffer.from(s, "utf8") for a batch: one byte array + offsets (what every batch entry point takes). */
export function pack(texts: string[]): { bytes: Uint8Array; offsets: Uint32Array } {
  const parts = texts.map((t) => enc.encode(t));
  const offsets = new Uint32Array(parts.length + 1);
  parts.forEach((p, i) => (offsets[i + 1] = offsets[i] + p.length));
  const bytes = new Uint8Array(offsets[parts.length] + 16);            // 16 bytes of slack: the device reads whole chunks
  parts.forEach((p, i) => bytes.set(p, offsets[i]));
  return { bytes, offsets };
}

export function initGpu(device?: number): void {
  addon.init(device);                                                  // throws without a usable B200: failMode decides (src/hooks.ts:232-241)
}

/** Drop-in for PatternRegistry (src/redaction/registry.ts:157-316): same constructor arguments, same findMatches result. */
export class GpuPatternRegistry {
  private readonly patterns: { id: string; category: string; regex: RegExp; replacementType: string }[];
  private readonly handle: unknown;

  constructor(patterns: { id: string; category: string; regex: RegExp; replacementType: string }[]) {
    const created = addon.createRuleset(
      patterns.map((p) => ({ source: p.regex.source, flags: p.regex.ignoreCase ? FLAG_ICASE : 0, category: CATEGORY[p.category] ?? 3 })),
    );
    // a pattern outside the supported subset never matches and is dropped, exactly like a failed compileCustomPattern (registry.ts:249-281)
    this.patterns = patterns.filter((_, i) => created.status[i] === 0);
    this.handle = addon.createRuleset(
      this.patterns.map((p) => ({ source: p.regex.source, flags: p.regex.ignoreCase ? FLAG_ICASE : 0, category: CATEGORY[p.category] ?? 3 })),
    ).handle;
  }

  /** findMatches for many strings at once; per string the same array PatternRegistry.findMatches returns. */
  findMatchesBatch(texts: string[]): { pattern: (typeof this.patterns)[number]; match: string; index: number; length: number }[][] {
    const { bytes, offsets } = pack(texts);
    const spans = addon.findMatchesBatch(this.handle, bytes, offsets);  // msg, rule, startByte, endByte, start16, end16 -- overlaps already resolved
    const out: ReturnType<GpuPatternRegistry["findMatchesBatch"]> = texts.map(() => []);
    for (let k = 0; k < spans.length; k += 6) {
      const [msg, rule, , , s16, e16] = spans.subarray(k, k + 6);
      out[msg].push({ pattern: this.patterns[rule], match: texts[msg].slice(s16, e16), index: s16, length: e16 - s16 });
    }
    return out;
  }

  findMatches(text: string) {
    return this.findMatchesBatch([text])[0];
  }
}

/** matchesAny (src/conditions/context.ts:9-25) for a whole policy set: every pattern is one rule, one scan per batch of texts. */
export class GpuRuleScanner {
  private readonly handle: unknown;
  private readonly ruleOfPattern: number[] = [];

  constructor(rules: (string | string[])[]) {
    const srcs: { source: string }[] = [];
    rules.forEach((pats, ri) => {
      for (const p of Array.isArray(pats) ? pats : [pats]) {
        // context.ts:15-17: a pattern that is not a valid RegExp falls back to includes() -- compiled as an escaped literal
        const source = addon.ruleCheck(p) === -4 /* CG_ERR_SYNTAX */ ? p.replace(/[.*+?^${}()|[\]\\\/-]/g, "\\$&") : p;
        srcs.push({ source });
        this.ruleOfPattern.push(ri);
      }
    });
    this.handle = addon.createRuleset(srcs).handle;
  }

  /** -> per text the sorted rule indices with a matching pattern */
  scan(texts: string[]): number[][] {
    const { bytes, offsets } = pack(texts);
    const { hits } = addon.scanBatch(this.handle, bytes, offsets);
    const out: Set<number>[] = texts.map(() => new Set());
    for (let k = 0; k < hits.length; k += 2) out[hits[k]].add(this.ruleOfPattern[hits[k + 1]]);
    return out.map((s) => [...s].sort((a, b) => a - b));
  }
}

/** sha256 (src/util.ts:77-79) for a batch of strings, hex digests. */
export function sha256Batch(texts: string[]): string[] {
  const parts = texts.map((t) => enc.encode(t));
  const offsets = new BigUint64Array(parts.length + 1);
  parts.forEach((p, i) => (offsets[i + 1] = offsets[i] + BigInt(p.length)));
  const bytes = new Uint8Array(Number(offsets[parts.length]) + 64);
  parts.forEach((p, i) => bytes.set(p, Number(offsets[i])));
  const d = addon.sha256Batch(bytes, offsets);
  return texts.map((_, i) => Buffer.from(d.subarray(32 * i, 32 * i + 32)).toString("hex"));
}

/** Proof-of-Guardrails: the day's audit JSONL (src/audit-trail.ts:151-179) as an append-only Merkle log.  flush() hands the
 * bytes it has just appended to the file to append(); root / frontier go next to the day file. */
export class AuditMerkleLog {
  private readonly log: unknown;
  constructor(restored?: { size: bigint; frontier: Uint8Array }) {
    this.log = restored ? addon.logRestore(restored.size, restored.frontier) : addon.logCreate(true);
  }
  append(flushedBytes: Uint8Array): bigint { return addon.logAppendJsonl(this.log, flushedBytes); }
  root(): string { return Buffer.from(addon.logRoot(this.log)).toString("hex"); }
  state(): { size: bigint; frontier: Uint8Array } { return { size: addon.logSize(this.log), frontier: addon.logFrontier(this.log) }; }
  inclusionProof(index: bigint): Uint8Array { return addon.logProof(this.log, index); }
  consistencyProof(firstSize: bigint): Uint8Array { return addon.logConsistency(this.log, firstSize); }
}
