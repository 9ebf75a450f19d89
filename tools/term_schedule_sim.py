"""Traffic model of a TERM-LEVEL (right-looking, operand-driven) schedule of the dot-product instructions of a
dependency group against the emitter's instruction-level chunks (round-1 plan "term-level register tiling"): every
accumulator still adds its terms in the reference's order (rounding unchanged), but a term is applied as soon as its
operands are final.  Loads fall, yet the accumulators that are open at a chunk boundary have to be spilled and
reloaded: the KKT factorisation of mpc_quadcopter gains at most 1.11x (LIFO wake-up order), the solve group 1.08x,
the residual group loses -- the working set of the sparse factor simply exceeds the register file.
    python tools/term_schedule_sim.py [workload]    -> profiles/r02_term_schedule_sim.txt"""
import sys, collections, time
import numpy as np
from tenscalc_b200 import problems, itable as it, emit_cuda
name = sys.argv[1] if len(sys.argv)>1 else "mpc_quadcopter"
gen,_ = problems.BUILDERS[name]()
pr = gen.problem; tab = pr.tab
ext = emit_cuda.external_values(pr)
CONST = (it.I_load, it.I_2)
DOT = {it.I_plus_minus_dot_div:2, it.I_minus_dot_div:1, it.I_plus_minus_dot:1, it.I_minus_dot:0}   # number of head operands

MODE='prio'
def micro_ops(g):
    """eager (operand-driven) term-level schedule of group g -> list of micro-ops
       ('term', consumer, a, b) / ('fin', consumer, heads) / ('op', instr, operands)"""
    members = pr.groups[g]; mset = set(members)
    terms = {}; heads = {}
    uses = collections.defaultdict(list)      # value -> [(consumer, term index or -1 for head/atomic)]
    for i in members:
        t = tab.types[i]; ops = [o for o in tab.operands[i]]
        if t in DOT:
            h = DOT[t]; heads[i] = ops[:h]; terms[i] = [(ops[k], ops[k+1]) for k in range(h, len(ops), 2)]
        elif t == it.I_sumprod and int(tab.params[i][0]) == 2 and len(ops) > 2:
            heads[i] = []; terms[i] = [(ops[k], ops[k+1]) for k in range(0, len(ops), 2)]
        else:
            heads[i] = ops; terms[i] = []
    final = set()
    def is_final(x): return x in final or x not in mset or tab.types[x] in CONST
    nxt = {i:0 for i in members}
    done = set()
    out = []
    waiting = collections.defaultdict(list)   # value -> consumers blocked on it
    def advance(c):
        # apply as many terms as possible, then finalise
        while nxt[c] < len(terms[c]):
            a,b = terms[c][nxt[c]]
            if not is_final(a): waiting[a].append(c); return
            if not is_final(b): waiting[b].append(c); return
            out.append(('term', c, a, b)); nxt[c] += 1
        for h in heads[c]:
            if not is_final(h): waiting[h].append(c); return
        if c in done: return
        done.add(c)
        out.append(('fin', c, tuple(heads[c])))
        final.add(c)
        _q.append(c)
    import heapq
    seeds_ = []
    order = {i:k for k,i in enumerate(members)}
    q = []
    class Q:
        def append(self, x):
            if MODE == "prio": heapq.heappush(q, (order[x], x))
            else: q.append((order[x], x))
    # every instruction registers at the start (it waits on its first missing operand); then a value that becomes
    # final wakes its waiters: the terms a new value feeds are applied at once (right-looking)
    globals()['_q'] = Q()
    # registration: nothing internal is final yet, so every instruction parks on its first internal operand; the
    # ones without internal operands are the seeds (processed through the same queue, in original order)
    internal = lambda x: x in mset and tab.types[x] not in CONST
    for i in members:
        first = None
        for a, b in terms[i]:
            if internal(a): first = a; break
            if internal(b): first = b; break
        if first is None:
            for h in heads[i]:
                if internal(h): first = h; break
        if first is None: seeds_.append(i)
        else: waiting[first].append(i)
    for i in seeds_:
        advance(i)
    while q:
        if MODE == "prio": _, x = heapq.heappop(q)
        elif MODE == "lifo": _, x = q.pop()
        else: _, x = q.pop(0)
        for c in sorted(waiting.pop(x, []), key=order.get):
            advance(c)
    assert len(done) == len(members), (len(done), len(members))
    return out

def chunk_cost(ops, g, max_loads=72, max_ops=1500):
    """greedy chunking of a micro-op stream: per chunk, distinct items loaded (values not produced in the chunk,
    accumulators opened before it) <= max_loads; stores: finals needed outside the chunk (ext or used later),
    accumulators still open at the end of the chunk"""
    members = set(pr.groups[g])
    # last use position of every value
    last = {}
    for p,op in enumerate(ops):
        if op[0]=='term': vs=(op[2],op[3])
        elif op[0]=='fin': vs=op[2]
        else: vs=op[2]
        for v in vs: last[v]=p
    nterm = collections.Counter(op[1] for op in ops if op[0]=='term')
    loads=stores=0; chunks=0
    p=0; n=len(ops)
    seen_terms = collections.Counter()
    while p<n:
        loaded=set(); made=set(); accs=set(); cnt=0; q=p
        while q<n:
            op=ops[q]
            if op[0]=='term': vs=[op[2],op[3]]
            else: vs=list(op[2])
            new=set(v for v in vs if v not in loaded and v not in made and tab.types[v] not in CONST)
            newacc = 0
            if op[0]=='term' and op[1] not in accs and seen_terms[op[1]]>0: newacc=1   # accumulator carried in
            if op[0]=='fin' and op[1] not in accs and nterm[op[1]]>0 and seen_terms[op[1]]>0: newacc=1
            if cnt and (len(loaded)+len(new)+newacc > max_loads or cnt>=max_ops): break
            loaded|=new
            if op[0]=='term':
                if newacc: loaded.add(('acc',op[1]))
                accs.add(op[1]); seen_terms[op[1]]+=1
            elif op[0]=='fin':
                if newacc: loaded.add(('acc',op[1]))
                accs.add(op[1]); made.add(op[1])
            cnt+=1; q+=1
        loads+=len(loaded)
        # stores
        for c in accs:
            if c in made:
                if c in ext or last.get(c,-1) >= q: stores+=1
            else:
                stores+=1     # open accumulator spilled
        chunks+=1; p=q
    return loads, stores, chunks

def baseline(g, max_loads=72):
    fs, st = emit_cuda.emit_group(pr, g, ext, 1536, max_loads, 1, 24)
    return st['loads'], st['stores']

groups = sorted(range(len(pr.groups)), key=lambda g:-len(pr.groups[g]))[:6]
for g in groups:
    t=time.time()
    b = baseline(g)
    for mode in ("prio","lifo","fifo"):
        globals()['MODE']=mode
        ops = micro_ops(g)
        r = chunk_cost(ops, g)
        print(f"   {mode}: {r[0]}+{r[1]}={r[0]+r[1]} ({r[2]} chunks) ratio {sum(b)/(r[0]+r[1]):.2f}")
    # in-order micro-ops (no rescheduling) for comparison
    print(f"group {g}: instr {len(pr.groups[g])} micro-ops {len(ops)}  baseline loads+stores {b[0]}+{b[1]}={sum(b)}  eager-term chunks: {r[0]}+{r[1]}={r[0]+r[1]} ({r[2]} chunks)  ratio {sum(b)/(r[0]+r[1]):.2f}  ({time.time()-t:.1f}s)")
