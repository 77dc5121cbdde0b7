import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np
import agrifly_b200 as agf
import orc_rappids as R
scn = agf.scenarios
N, K = 256, 512
for variant in ('easy','hard'):
    kw = {} if variant=='easy' else dict(speed_max=4.5, acc_max=3.0, box_depth=(1.0,3.0), n_boxes=(2,4))
    pop = scn.rappids_population(N, seed=11, **kw)
    imgs = scn.rappids_render(pop['row_bg'], pop['boxes'], 320)
    ocfg = R.default_cfg(max_pyramids=32)
    port = R.Planner('port-shared')
    # candidates from the reference-equivalent mt19937 sampler
    cands = np.zeros((N,K,4)); exp=[]
    t=time.time()
    for i in range(N):
        r = port.plan(ocfg, imgs[i], pop['vel0'][i], pop['acc0'][i], pop['grav'][i], n=K, seed=i)
        cands[i] = r['candidates']; exp.append(r)
    print(variant, 'oracle time', time.time()-t)
    cfg = agf.rappids_cfg(math=agf.abi.MATH_PARITY)
    with agf.Rappids(cfg, N, K) as pl:
        pl.render_scenes(pop['row_bg'], pop['boxes'])
        assert np.array_equal(pl.get_images(), imgs), 'raster mismatch'
        pl.set_states(pop['vel0'], pop['acc0'], pop['grav'])
        pl.set_candidates(cands)
        pl.plan(); pl.sync()
        res = pl.results(); fl = pl.candidate_flags(); py = pl.pyramids()
        print('kernel ms', pl.plan_kernel_time(), pl.stats())
    bad = 0
    for i in range(N):
        e = exp[i]
        ok = True
        for k in ('found','best_index','n_generated','n_cost_checks','n_collision_checks','n_velocity_checks','n_collision_free','n_pyramids'):
            if res[i][k] != e[k]: ok=False; print(i,k,res[i][k],e[k])
        if not np.array_equal(fl[i], e['results']): ok=False; print(i,'flags differ', np.nonzero(fl[i]!=e['results'])[0][:5])
        if e['found']:
            if not np.array_equal(res[i]['best_coeffs'], e['best_coeffs']): ok=False; print(i,'coeffs differ', np.abs(res[i]['best_coeffs']-e['best_coeffs']).max())
            if res[i]['best_cost'] != e['best_cost'] or res[i]['best_tf'] != e['best_tf']: ok=False; print(i,'cost/tf')
        np_ = e['n_pyramids']
        if not np.array_equal(py[i,:np_], e['pyramids'][:np_]): ok=False; print(i,'pyramids differ', np.abs(py[i,:np_]-e['pyramids'][:np_]).max())
        bad += (not ok)
    print(variant, 'mismatching vehicles', bad, 'of', N, '| found', int(res['found'].sum()), 'mean pyramids', res['n_pyramids'].mean(), 'mean cost checks', res['n_cost_checks'].mean(), 'vel ok', res['n_velocity_checks'].mean(), 'free', res['n_collision_free'].mean())
    # fast variant
    cfg = agf.rappids_cfg(math=agf.abi.MATH_FAST)
    with agf.Rappids(cfg, N, K) as pl:
        pl.render_scenes(pop['row_bg'], pop['boxes']); pl.set_states(pop['vel0'], pop['acc0'], pop['grav']); pl.set_candidates(cands)
        pl.plan(); pl.sync(); r2 = pl.results(); f2 = pl.candidate_flags()
        print('fast: kernel ms', pl.plan_kernel_time(), 'same best idx', int((r2['best_index']==res['best_index']).sum()), 'same flags', int((f2==fl).all(axis=1).sum()))
