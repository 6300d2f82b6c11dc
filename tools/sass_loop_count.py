"""Static instruction mix of the LARGEST loop of a kernel (the per-step loop of the thread-per-series kernels) from
`cuobjdump -sass -fun <mangled name> <object>` output:

    cuobjdump -sass -fun _ZN5physs22seq_smooth_pipe_kernelILi4ELi4ELi0ELi0ELb0EEEvNS_13SeqSmoothArgsE \
        physs_gp_b200/csrc/_obj/physs_seq_d4s4m.o > /tmp/k.sass && python tools/sass_loop_count.py /tmp/k.sass

Used for the FP64-issue accounting in DESIGN.md section 3 (profiles/sass_r02_d4_loop_counts.json)."""
import re,sys,collections
for fn in sys.argv[1:]:
    ins=[]
    for line in open(fn):
        m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',line)
        if m: ins.append((int(m.group(1),16),m.group(2)))
    # find backward branches
    best=None
    for i,(addr,txt) in enumerate(ins):
        m=re.search(r'\bBRA\b.*?0x([0-9a-f]+)',txt)
        if m:
            tgt=int(m.group(1),16)
            if tgt<addr:
                n=sum(1 for a,_ in ins if tgt<=a<=addr)
                if best is None or n>best[0]: best=(n,tgt,addr)
    n,tgt,addr=best
    body=[t for a,t in ins if tgt<=a<=addr]
    c=collections.Counter()
    for t in body:
        t=re.sub(r'^@!?U?P\d+\s+','',t)
        op=t.split()[0].split('.')[0]
        c[op]+=1
    fp64=sum(v for k,v in c.items() if k in('DFMA','DADD','DMUL','DSETP','DMNMX'))
    print(fn,'loop instrs',n,'FP64-pipe',fp64,'MUFU',c['MUFU'], 'LDS',c['LDS'],'STS',c['STS'],'LDG',c['LDG'],'STG',c['STG'],'LDGSTS',c['LDGSTS'],'LDL',c['LDL'],'STL',c['STL'])
    print('   top:',c.most_common(14))
