import sys, json, csv, collections
for f in sys.argv[1:]:
    if f.endswith('.csv'):
        lines=[l for l in open(f) if not l.startswith('==')]
        agg=collections.defaultdict(list)
        for row in csv.DictReader(lines):
            if row.get('Metric Name')=='gpu__time_duration.sum':
                v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
                v = v/1e6 if u=='ns' else v/1e3 if u in ('us','usecond') else v*1e3 if u=='s' else v
                agg[row['Kernel Name'][:58]].append(v)
        print(f)
        for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:8]:
            print("   %-60s n=%3d  avg %.3f ms  total %.2f ms"%(k,len(v),sum(v)/len(v),sum(v)))
        continue
    try:
        l=[x for x in open(f) if x.startswith('{')][-1]; d=json.loads(l)
        print("%-32s iter_ms %.3f spmv_ms %.3f frac %.3f e2e_s %.4f launches %d sm_mhz %s energy %s"%(f.split('/')[-1], d['ms_per_step'], d['spmv_ms'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'] if d.get('clocks') else None, d.get('energy')))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-600:])
