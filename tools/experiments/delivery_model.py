"""Discrete-event model of the host delivery of rwrt_real2d_host: GPU stream, copy stream, host
writer threads and the launching thread, per time window.  Used to choose the routing rule of hx_on_window (a window
goes straight into the caller's pinned planes when `backlog >= P` host windows are still unexpanded; decisions are
taken when the copy of window w - `throttle` has finished) before spending GPU time on it.

    python tools/experiments/delivery_model.py

a = link time of a host-route window (packed rows), a direct window costs 2a; c = time the writer threads need per
window; rates below are per half-GB window of C2 on the boxes of profiles/r1_v16_e2e_sweep.txt.  The rule that was
built is P = 4 (= all staging buffers busy), throttle = 2: never worse than the pure host route in the model, 317 ->
277 ms on the fast-link box (measured there: 298 -> 278 ms)."""
def sim(nwin, a, c, kern, rule, P=3, stages=4, throttle=2, verbose=False):
    # times in ms. a: link time of an hx window; direct = 2a. c: CPU time per hx window. kern: GPU time per window
    t_main = 0.0
    gpu_free = 0.0
    link_free = 0.0
    cpu_free = 0.0
    copied = []          # completion time of each window's copy
    hx_done = []         # completion times of hx windows' expansion (in order)
    hx_copy_done = []
    routes = []
    last_direct = False
    for w in range(nwin):
        # launch integrate(w): GPU
        k_start = max(gpu_free, t_main)
        k_end = k_start + kern
        gpu_free = k_end
        # decision point: throttle on copied[w-throttle]
        if w - throttle >= 0:
            t_main = max(t_main, copied[w - throttle])
        # pending at t_main
        pending = sum(1 for i, d in enumerate(hx_done) if d > t_main)
        if rule == 'hx': direct = False
        elif rule == 'direct': direct = True
        elif rule == 'auto': direct = (pending >= P) and not last_direct
        elif rule == 'auto2': direct = (pending >= P)
        routes.append(direct)
        last_direct = direct
        if not direct:
            nh = len(hx_done)
            if nh - stages >= 0:
                t_main = max(t_main, hx_done[nh - stages])   # staging reuse
            c_start = max(link_free, gpu_free)     # pack negligible
            c_end = c_start + a
            link_free = c_end
            copied.append(c_end)
            e_start = max(cpu_free, c_end)
            e_end = e_start + c
            cpu_free = e_end
            hx_done.append(e_end)
        else:
            c_start = max(link_free, gpu_free)
            c_end = c_start + 2 * a
            link_free = c_end
            copied.append(c_end)
    total = max(link_free, cpu_free)
    return total, sum(routes)
for name, (a, c) in {'fast link box A': (5.0, 6.6), 'mid link box C': (6.0, 6.6), 'slow link box B': (9.0, 6.6), 'N=2 box (both inflated)': (8.0, 14.0)}.items():
    print(name)
    for rule, P in (('hx', 0), ('direct', 0), ('auto', 2), ('auto', 3), ('auto2', 3), ('auto2', 4), ('auto',4)):
        for thr in (1, 2):
            tot, nd = sim(47, a, c, 1.5, rule, P=P, throttle=thr)
            print(f"   {rule:7s} P={P} throttle={thr}: {tot:7.1f} ms, {nd} direct windows")
