import json,sys
d=json.load(open(sys.argv[1]))
for k,v in d.items():
    c=v['certified']; print(k, 'cand %.3f exact %.3f sum %.3f'%(c['candidates_ms'],c['exactness_ms'],c['candidates_ms']+c['exactness_ms']), 'sat',c['stats']['saturated_queries'],'pairs',c['stats']['pairs_rescored'],'emit',c['stats']['second_pass_pairs_emitted'],'viol',c['stats']['certified_bound_violations'],'diff',c['indices_differing_from_exhaustive'], 'ratio %.3f'%v['certified_over_fixed'])
