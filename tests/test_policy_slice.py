"""The policy-evaluation slice and the response gate (plugin.PolicyEvaluator / plugin.ResponseGate), restated from the
reference's own tests: gov/test/policy-evaluator.test.ts:59-304, gov/test/conditions/tool.test.ts:60-145,
gov/test/conditions/context.test.ts:31-45,107-110, gov/test/response-gate.test.ts:13-180,217-260.

Run twice: on the CPU tier with an oracle-backed scanner (host logic only), on the GPU tier with the product's
RuleScanner (every regex test goes through the kernels, one batch per call)."""
import numpy as np
import pytest

from vainplex_openclaw_b200 import plugin as P


def make_ctx(**over):
    # policy-evaluator.test.ts:8-55 (makeCtx)
    ctx = {"hook": "before_tool_call", "agentId": "forge", "sessionKey": "agent:main:subagent:forge:abc",
           "trust": {"agent": {"tier": "trusted"}, "session": {"tier": "standard"}},
           "toolName": "exec", "toolParams": {"command": "docker rm container-x"}, "channel": "matrix",
           "messageContent": "Hello world", "conversationContext": ["We discussed JIRA-1234 yesterday", "and other things"]}
    ctx.update(over)
    return ctx


def tool_rule(effect, rid="r1", **kw):
    return {"id": rid, "conditions": [{"type": "tool", "name": "exec"}], "effect": effect, **kw}


@pytest.fixture(params=["oracle", pytest.param("device", marks=pytest.mark.gpu)])
def scanner_cls(request, oracle):
    if request.param == "device":
        from vainplex_openclaw_b200 import _native as N
        N.init(-1)
        return None                     # the product's RuleScanner

    class OracleScanner:                # same interface; oracle.matches_any is the oracle's restatement of context.ts:9-25
        def __init__(self, rules, logger=None):
            self.rules = [([pats] if isinstance(pats, str) else list(pats)) for pats in rules]
            self.cache = {}

        def scan(self, texts):
            return [[ri for ri, pats in enumerate(self.rules) if oracle.matches_any(pats, [t], self.cache)] for t in texts]
    return OracleScanner


def ev(policies, ctx=None, scanner_cls=None):
    return P.PolicyEvaluator(policies, scanner_cls=scanner_cls).evaluate(ctx or make_ctx())


def test_policy_evaluator_reference_cases(scanner_cls):
    S = scanner_cls
    r = ev([], scanner_cls=S)                                                   # :60-65
    assert r["action"] == "allow" and r["matches"] == [] and r["reason"] == "No matching policies"
    r = ev([{"id": "p1", "scope": {}, "rules": [tool_rule({"action": "deny", "reason": "No exec allowed"})]}], scanner_cls=S)   # :67-86
    assert (r["action"], r["reason"], len(r["matches"])) == ("deny", "No exec allowed", 1)
    allow_p = {"id": "allow-p", "scope": {}, "rules": [tool_rule({"action": "allow"})]}
    deny_p = {"id": "deny-p", "scope": {}, "rules": [tool_rule({"action": "deny", "reason": "Denied"})]}
    assert ev([allow_p, deny_p], scanner_cls=S)["action"] == "deny"            # :88-118 deny wins
    low = {"id": "low", "scope": {}, "priority": 1, "rules": [tool_rule({"action": "allow"})]}
    high = {"id": "high", "scope": {}, "priority": 10, "rules": [tool_rule({"action": "deny", "reason": "High priority deny"})]}
    r = ev([low, high], scanner_cls=S)                                          # :120-153
    assert r["action"] == "deny" and [m["policyId"] for m in r["matches"]] == ["high", "low"]
    assert ev([{"id": "p1", "scope": {"excludeAgents": ["forge"]}, "rules": [tool_rule({"action": "deny", "reason": "Denied"})]}], scanner_cls=S)["action"] == "allow"   # :155-172
    assert ev([{"id": "p1", "scope": {"channels": ["telegram"]}, "rules": [tool_rule({"action": "deny", "reason": "Denied"})]}], scanner_cls=S)["action"] == "allow"     # :174-192
    r = ev([{"id": "p1", "scope": {}, "rules": [tool_rule({"action": "allow"}, minTrust="trusted")]}], scanner_cls=S)      # :194-214
    assert r["action"] == "allow" and r["matches"] == []
    assert ev([{"id": "p1", "scope": {}, "rules": [tool_rule({"action": "deny", "reason": "Too new"}, maxTrust="restricted")]}], scanner_cls=S)["action"] == "allow"   # :216-235
    r = ev([{"id": "p1", "scope": {}, "rules": [tool_rule({"action": "audit", "level": "verbose"})]}], scanner_cls=S)       # :237-255
    assert (r["action"], r["reason"]) == ("allow", "Allowed with audit logging")
    r = ev([{"id": "p1", "scope": {}, "rules": [{"id": "r1", "conditions": [{"type": "tool", "name": "exec"}, {"type": "agent", "id": "cerberus"}],
                                                  "effect": {"action": "deny", "reason": "Denied"}}]}], scanner_cls=S)        # :257-277 AND
    assert r["action"] == "allow"
    r = ev([{"id": "p1", "scope": {}, "rules": [tool_rule({"action": "allow"}), tool_rule({"action": "deny", "reason": "Should not reach"}, rid="r2")]}], scanner_cls=S)   # :279-304
    assert r["action"] == "allow" and r["matches"][0]["ruleId"] == "r1" and r["reason"] == "Allowed by governance policy"
    r = ev([{"id": "p1", "scope": {}, "controls": ["A.8.3"], "rules": [tool_rule({"action": "allow"})]}], scanner_cls=S)    # :306-345
    assert r["matches"][0]["controls"] == ["A.8.3"]
    # policy-evaluator.ts:55-68: 2fa sits between deny and audit; specificity breaks priority ties (:28-42)
    twofa = {"id": "t", "scope": {}, "rules": [tool_rule({"action": "2fa"})]}
    audit = {"id": "a", "scope": {"agents": ["forge"], "channels": ["matrix"]}, "rules": [tool_rule({"action": "audit"})]}
    r = ev([twofa, audit], scanner_cls=S)
    assert (r["action"], r["reason"]) == ("2fa", "Requires 2FA approval") and [m["policyId"] for m in r["matches"]] == ["a", "t"]
    assert ev([twofa, audit, deny_p], scanner_cls=S)["action"] == "deny"


def test_conditions_on_the_scan_path(scanner_cls):
    S = scanner_cls

    def cond_true(cond, **over):
        pol = [{"id": "p", "scope": {}, "rules": [{"id": "r", "conditions": [cond], "effect": {"action": "deny", "reason": "x"}}]}]
        return ev(pol, make_ctx(**over), scanner_cls=S)["action"] == "deny"
    # conditions/tool.test.ts:70-80,118-145
    c = {"type": "tool", "name": "exec", "params": {"command": {"matches": "git push.*(main|master)"}}}
    assert cond_true(c, toolParams={"command": "git push origin main"}) and not cond_true(c, toolParams={"command": "git push origin dev"})
    assert cond_true({"type": "tool", "params": {"command": {"matches": "docker.*"}}})
    assert not cond_true({"type": "tool", "params": {"command": {"matches": "[invalid"}}})
    assert not cond_true({"type": "tool", "params": {"count": {"matches": "\\d+"}}}, toolParams={"count": 42})
    assert cond_true({"type": "tool", "params": {"elevated": {"equals": True}}}, toolParams={"elevated": True})
    assert not cond_true({"type": "tool", "params": {"elevated": {"equals": True}}}, toolParams={"elevated": False})
    assert cond_true({"type": "tool", "params": {"command": {"startsWith": "sudo"}}}, toolParams={"command": "sudo rm -rf"})
    assert cond_true({"type": "tool", "params": {"host": {"in": ["sandbox", "gateway"]}}}, toolParams={"host": "sandbox"})
    assert not cond_true({"type": "tool", "params": {"host": {"in": ["sandbox", "gateway"]}}}, toolParams={"host": "node"})
    assert cond_true({"type": "tool"}) and not cond_true({"type": "tool", "name": "exec"}, toolName=None)
    assert not cond_true({"type": "tool", "params": {"command": {"contains": "test"}}}, toolParams=None)
    assert cond_true({"type": "tool", "name": "ex*"}) and not cond_true({"type": "tool", "name": "e?"})
    # conditions/context.test.ts:32-45,107-110
    assert cond_true({"type": "context", "conversationContains": "JIRA-\\d+"})
    assert not cond_true({"type": "context", "conversationContains": "TICKET-\\d+"})
    assert not cond_true({"type": "context", "conversationContains": "JIRA"}, conversationContext=[])
    assert cond_true({"type": "context", "conversationContains": ["NOPE", "JIRA-\\d+"]})
    assert cond_true({"type": "context", "messageContains": "Hello"}) and not cond_true({"type": "context", "messageContains": "Hello"}, messageContent="")
    assert cond_true({"type": "context", "messageContains": "a(b"}, messageContent="xx a(b yy")       # context.ts:15-17 includes() fallback
    # conditions/simple.ts:124-150 any / not
    assert cond_true({"type": "any", "conditions": [{"type": "context", "messageContains": "zzz"}, {"type": "context", "conversationContains": "JIRA"}]})
    assert cond_true({"type": "not", "condition": {"type": "context", "messageContains": "zzz"}})
    assert not cond_true({"type": "not", "condition": {"type": "context", "messageContains": "Hel+o"}})
    # a batch: one scan for all contexts
    pe = P.PolicyEvaluator([{"id": "p", "scope": {}, "rules": [{"id": "r", "conditions": [{"type": "context", "messageContains": "secret-\\d{3}"}], "effect": {"action": "deny", "reason": "leak"}}]}], scanner_cls=S)
    res = pe.evaluate_batch([make_ctx(messageContent="x secret-%03d y" % i if i % 3 == 0 else "nothing %d" % i) for i in range(60)])
    assert [r["action"] for r in res] == ["deny" if i % 3 == 0 else "allow" for i in range(60)]


def test_response_gate_reference_cases(scanner_cls):
    S = scanner_cls

    def gate(rules, **kw):
        return P.ResponseGate({"enabled": True, "rules": rules, **kw}, scanner_cls=S)
    assert P.ResponseGate({"enabled": False, "rules": [{"validators": [{"type": "mustMatch", "pattern": "x"}]}]}, scanner_cls=S).validate("y", "main", [])["passed"]   # :13-22
    g = gate([{"validators": [{"type": "requiredTools", "tools": ["web_search"]}]}])
    assert g.validate("a", "main", [{"toolName": "web_search", "output": "ok"}])["passed"] and not g.validate("a", "main", [])["passed"]          # :24-42
    r = gate([{"validators": [{"type": "requiredTools", "tools": ["a", "b"], "message": "Must verify facts first!"}]}]).validate("claim", "main", [{"toolName": "a", "output": ""}])
    assert not r["passed"] and r["reasons"][0] == "Must verify facts first!" and r["failedValidators"] == ["requiredTools:a,b"]     # :44-65
    assert gate([{"validators": [{"type": "mustMatch", "pattern": "\\d+"}]}]).validate("result: 42", "main", [])["passed"]          # :69-74
    r = gate([{"validators": [{"type": "mustMatch", "pattern": "^VERIFIED:"}]}]).validate("just a normal response", "main", [])      # :76-84
    assert not r["passed"] and "mustMatch" in r["failedValidators"][0]
    assert gate([{"validators": [{"type": "mustNotMatch", "pattern": "password:\\s*\\S+"}]}]).validate("safe content", "main", [])["passed"]   # :87-92
    r = gate([{"validators": [{"type": "mustNotMatch", "pattern": "I don't know"}]}]).validate("I don't know the answer", "main", [])  # :94-102
    assert not r["passed"] and "mustNotMatch" in r["failedValidators"][0]
    g = gate([{"agentId": "cerberus", "validators": [{"type": "mustNotMatch", "pattern": ".*"}]}])                                  # :105-113
    assert g.validate("anything", "main", [])["passed"] and not g.validate("anything", "cerberus", [])["passed"]
    g = gate([{"agentId": ["forge", "atlas"], "validators": [{"type": "mustMatch", "pattern": "^OK"}]}])                            # :115-122
    assert g.validate("OK done", "forge", [])["passed"] and not g.validate("nope", "forge", [])["passed"] and g.validate("nope", "main", [])["passed"]
    g = gate([{"validators": [{"type": "requiredTools", "tools": ["web_search"]}, {"type": "mustNotMatch", "pattern": "placeholder"}]}])   # :135-158
    ok = [{"toolName": "web_search", "output": "ok"}]
    assert not g.validate("real answer", "main", [])["passed"] and not g.validate("placeholder text", "main", ok)["passed"] and g.validate("real answer", "main", ok)["passed"]
    for bad, typ in (("[", "mustMatch"), ("(?P<bad)", "mustNotMatch")):                                                              # :161-178
        r = gate([{"validators": [{"type": typ, "pattern": bad}]}]).validate("anything", "main", [])
        assert not r["passed"] and "invalid regex" in r["reasons"][0]
    r = gate([{"validators": [{"type": "mustNotMatch", "pattern": "BLOCKED"}]}], fallbackTemplate="{agent}: {validators} -- {reasons}").validate("BLOCKED", "forge", [])   # :236-245
    assert r["fallbackMessage"] == "forge: mustNotMatch:BLOCKED -- Response Gate: content matches forbidden pattern /BLOCKED/"
    assert "fallbackMessage" not in gate([{"validators": [{"type": "mustNotMatch", "pattern": "BLOCKED"}]}], fallbackMessage="x").validate("fine", "forge", [])   # :257-268
    res = gate([{"validators": [{"type": "mustNotMatch", "pattern": "token-[a-f0-9]{8}"}]}]).validate_batch([("t token-%08x" % i if i % 2 else "clean", "main", []) for i in range(40)])
    assert [r["passed"] for r in res] == [i % 2 == 0 for i in range(40)]
