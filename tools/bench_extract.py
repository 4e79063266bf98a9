import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
        print(f, 'value',round(d['value']),'single',round(d['single_session']['value']),'e2e',round(d['e2e']['value'],1),'rans',round(d['rans']['ms_longest_stream'],3))
    except Exception as e: print(f,'ERR',e)
