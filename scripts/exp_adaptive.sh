# sweep of the work-queue chunking knobs (run on the GPU box)
for IP in ${IPS:-2}; do for K in ${KS:-10 20 40 60}; do
VP_QUEUE_ITEMS_PER_CTA=$IP timeout 300 python bench.py --steps $K --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('items_per_cta=$IP K=$K value', round(d['value']), 'frac', round(d['roofline']['frac'],3), 'evals', d['config']['evals_per_fit_mean'], 'region_us', round(d['roofline']['region_us']), 'e2e', round(d['e2e']['value']))"
done; done
