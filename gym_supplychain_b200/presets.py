"""Topology presets: the constructor defaults of the reference's SupplyChainEnv subclasses, as data.

Each function takes the same kwargs as the reference class and returns `(nodes_info, env_kwargs)`
ready for `FlatSpec(nodes_info, **env_kwargs)`:

  two_per_stage ............ SupplyChain2perStageEnv            supplychain_2perstage_env.py:14-64
  two_per_stage_seasonal ... SupplyChain2perStageSeasonalEnv    supplychain_2perstage_env.py:67-97
  n_per_stage .............. SupplyChainNPerStage               supplychain_Nperstage_env.py:37-131
  multi_product (+variants)  SupplyChainMultiProduct*           supplychain_multiproduct_env.py:56-273
  main_demo_chain .......... the 2-2-2-2 chain of supplychain_env.py:842-868 (default for supplychain-v0)

tests/test_host_cpu.py (test_presets_reproduce_reference_nodes_info) pins every one of them against the
nodes_info the reference classes build (captured into tests/golden/*.npz).
"""

_ENV_KEYS = ("num_products", "unmet_demand_cost", "exceeded_stock_capacity_cost", "exceeded_process_capacity_cost",
             "exceeded_ship_capacity_cost", "processing_ratio", "demand_range", "demand_std", "demand_sen_peaks",
             "avg_demand_range", "total_time_steps", "stochastic_leadtimes", "avg_leadtime", "max_leadtime", "seed",
             "build_info", "demand_perturb_norm")


def _env_kwargs(values, **extra):
    kw = {k: values[k] for k in _ENV_KEYS}
    kw.update(extra)
    return kw


def _layered(names, echelons, dest_costs, ship_caps):
    """echelons: 4 lists of per-node dicts (without destinations); wires echelon k to k+1."""
    info = {}
    for k in range(4):
        for j, node in enumerate(echelons[k]):
            node = dict(node)
            if k < 3:
                node["destinations"] = list(names[k + 1])
                node["dest_costs"] = dest_costs[k]
                node["ship_capacity"] = ship_caps[k]
            else:
                node["last_level"] = True
            info[names[k][j]] = node
    return info


def two_per_stage(num_products=1, initial_stocks=[0] * 8, initial_supply=[[[60, 60]]] * 2,
                  initial_shipments=[[[60, 60]]] * 2 + [[[20, 20]]] * 4, supply_capacities=[120, 150],
                  processing_capacities=[300, 300], stock_capacities=[200, 300] * 4, ship_capacity=300,
                  processing_ratio=3, processing_costs=[12, 10], stock_costs=[1] * 8, supply_costs=[6, 4], dest_cost=2,
                  unmet_demand_cost=216, exceeded_stock_capacity_cost=10, exceeded_process_capacity_cost=10,
                  exceeded_ship_capacity_cost=10, demand_range=(10, 20), demand_std=None, demand_sen_peaks=None,
                  avg_demand_range=None, stochastic_leadtimes=False, avg_leadtime=2, max_leadtime=2,
                  total_time_steps=360, seed=None, build_info=False, demand_perturb_norm=False):
    if not initial_stocks:
        initial_stocks = [0] * 8
    names = [["Supplier1", "Supplier2"], ["Factory1", "Factory2"], ["WholeSaler1", "WholeSaler2"],
             ["Retailer1", "Retailer2"]]

    def base(i):
        return {"initial_stock": initial_stocks[i], "stock_capacity": stock_capacities[i], "stock_cost": stock_costs[i]}
    ech = [[], [], [], []]
    for i in range(2):
        ech[0].append(dict(base(i), initial_supply=initial_supply[i], supply_capacity=supply_capacities[i],
                           supply_cost=supply_costs[i]))
        ech[1].append(dict(base(2 + i), initial_shipments=initial_shipments[i],
                           processing_capacity=processing_capacities[i], processing_cost=processing_costs[i]))
        ech[2].append(dict(base(4 + i), initial_shipments=initial_shipments[2 + i]))
        ech[3].append(dict(base(6 + i), initial_shipments=initial_shipments[4 + i]))
    dc = [[dest_cost] * 2] * num_products
    info = _layered(names, ech, [dc] * 3, [[ship_capacity] * 2] * 3)
    return info, _env_kwargs(locals())


def two_per_stage_seasonal(num_products=1, initial_stocks=[800] * 8, initial_supply=[[[600, 600]], [[840, 840]]],
                           initial_shipments=[[[600, 600]], [[840, 840]]] + [[[240, 240]]] * 4,
                           supply_capacities=[600, 840], processing_capacities=[840, 960],
                           stock_capacities=[1600, 1800] * 4, ship_capacity=1800, processing_ratio=3,
                           processing_costs=[12, 10], stock_costs=[1] * 8, supply_costs=[6, 4], dest_cost=2,
                           unmet_demand_cost=216, exceeded_stock_capacity_cost=10, exceeded_process_capacity_cost=10,
                           exceeded_ship_capacity_cost=10, demand_range=(0, 400), demand_std=10, demand_sen_peaks=4,
                           avg_demand_range=(150, 250), stochastic_leadtimes=False, avg_leadtime=2, max_leadtime=2,
                           total_time_steps=360, seed=None, build_info=False, check_actions=False,
                           demand_perturb_norm=True):
    kw = dict(locals())
    kw.pop("check_actions")
    return two_per_stage(**kw)


def n_per_stage(nodes_per_echelon=3, num_products=2, initial_stocks=None, stock_capacities=None, stock_costs=1,
                initial_supply=None, supply_capacities=None, supply_costs=None, dest_cost=None, ship_capacity=None,
                initial_shipments=None, processing_capacities=None, processing_costs=None, processing_ratio=3,
                unmet_demand_cost=216, exceeded_stock_capacity_cost=10, exceeded_process_capacity_cost=10,
                exceeded_ship_capacity_cost=10, demand_range=(0, 400), demand_std=None, demand_sen_peaks=None,
                avg_demand_range=None, demand_perturb_norm=False, stochastic_leadtimes=False, avg_leadtime=2,
                max_leadtime=2, total_time_steps=360, seed=None, build_info=False):
    P = num_products
    if isinstance(nodes_per_echelon, int):
        nodes_per_echelon = [nodes_per_echelon] * 4
    s, f, w, r = nodes_per_echelon
    keys = ("suppliers", "factories", "wholesalers", "retailers")
    cnt = dict(zip(keys, (s, f, w, r)))
    if not stock_capacities:
        stock_capacities = {k: [[6400 if k == "factories" else 1600] * P] * cnt[k] for k in keys}
    if not initial_stocks:
        initial_stocks = {k: [[800] * P] * cnt[k] for k in keys}
    if not initial_supply:
        initial_supply = [[[600] * avg_leadtime] * P] * s
    if not supply_capacities:
        supply_capacities = [[600] * P] * s
    if not supply_costs:
        supply_costs = [[6] * P] * s
    if not dest_cost:
        dest_cost = {"suppliers": [[2] * f] * P, "factories": [[2] * w] * P, "wholesalers": [[2] * r] * P}
    if not ship_capacity:
        ship_capacity = {"suppliers": [500 * P] * f, "factories": [500 * P] * w, "wholesalers": [500 * P] * r}
    if not initial_shipments:
        initial_shipments = {"factories": [[[600] * avg_leadtime] * P] * f,
                             "wholesalers": [[[240] * avg_leadtime] * P] * w,
                             "retailers": [[[240] * avg_leadtime] * P] * r}
    if not processing_capacities:
        processing_capacities = [840 * P] * f
    if not processing_costs:
        processing_costs = [[12] * P] * f
    names = [["Supplier%d" % i for i in range(s)], ["Factory%d" % i for i in range(f)],
             ["Wholesal%d" % i for i in range(w)], ["Retailer%d" % i for i in range(r)]]

    def base(k, i):
        return {"initial_stock": initial_stocks[k][i], "stock_capacity": stock_capacities[k][i], "stock_cost": stock_costs}
    ech = [[dict(base("suppliers", i), initial_supply=initial_supply[i], supply_capacity=supply_capacities[i],
                 supply_cost=supply_costs[i]) for i in range(s)],
           [dict(base("factories", i), initial_shipments=initial_shipments["factories"][i],
                 processing_capacity=processing_capacities[i], processing_cost=processing_costs[i]) for i in range(f)],
           [dict(base("wholesalers", i), initial_shipments=initial_shipments["wholesalers"][i]) for i in range(w)],
           [dict(base("retailers", i), initial_shipments=initial_shipments["retailers"][i]) for i in range(r)]]
    info = _layered(names, ech, [dest_cost[k] for k in keys[:3]], [ship_capacity[k] for k in keys[:3]])
    return info, _env_kwargs(locals())


def multi_product(demand_config_by_product=False, num_products=2, initial_stocks=None, stock_capacities=None,
                  stock_costs=1, initial_supply=None, supply_capacities=None, supply_costs=None, dest_cost=None,
                  ship_capacity=None, initial_shipments=None, processing_capacities=None, processing_costs=None,
                  processing_ratio=3, unmet_demand_cost=216, exceeded_stock_capacity_cost=10,
                  exceeded_process_capacity_cost=10, exceeded_ship_capacity_cost=10, demand_range=(0, 400),
                  demand_std=None, demand_sen_peaks=None, avg_demand_range=None, demand_perturb_norm=False,
                  stochastic_leadtimes=False, avg_leadtime=2, max_leadtime=2, total_time_steps=360, seed=None,
                  build_info=False):
    P = num_products
    if not stock_capacities:
        stock_capacities = [[c] * P for c in (1600, 1800, 6400, 7200, 1600, 1800, 1600, 1800)]
    if not initial_stocks:
        initial_stocks = [[800] * P] * 8
    if not initial_supply:
        initial_supply = [[[600] * avg_leadtime] * P, [[840] * avg_leadtime] * P]
    if not supply_capacities:
        supply_capacities = [[600] * P, [840] * P]
    if not supply_costs:
        supply_costs = [[6] * P, [4] * P]
    if not dest_cost:
        dest_cost = [[2] * 2] * P
    if not ship_capacity:
        ship_capacity = [500 * P, 500 * P]
    if not initial_shipments:
        initial_shipments = [[[600] * avg_leadtime] * P, [[840] * avg_leadtime] * P] + [[[240] * avg_leadtime] * P] * 4
    if not processing_capacities:
        processing_capacities = [840 * P, 960 * P]
    if not processing_costs:
        processing_costs = [[12] * P, [10] * P]
    names = [["Supplier1", "Supplier2"], ["Factory1", "Factory2"], ["Wholesal1", "Wholesal2"], ["Retailer1", "Retailer2"]]

    def base(i):
        return {"initial_stock": initial_stocks[i], "stock_capacity": stock_capacities[i], "stock_cost": stock_costs}
    ech = [[dict(base(i), initial_supply=initial_supply[i], supply_capacity=supply_capacities[i],
                 supply_cost=supply_costs[i]) for i in range(2)],
           [dict(base(2 + i), initial_shipments=initial_shipments[i], processing_capacity=processing_capacities[i],
                 processing_cost=processing_costs[i]) for i in range(2)],
           [dict(base(4 + i), initial_shipments=initial_shipments[2 + i]) for i in range(2)],
           [dict(base(6 + i), initial_shipments=initial_shipments[4 + i]) for i in range(2)]]
    info = _layered(names, ech, [dest_cost] * 3, [ship_capacity] * 3)
    return info, _env_kwargs(locals(), demand_config_by_product=demand_config_by_product)


def _increasing_costs(P):
    return dict(supply_costs=[[6 * (i + 1) for i in range(P)], [4 * (i + 1) for i in range(P)]],
                dest_cost=[[2 * (i + 1)] * 2 for i in range(P)],
                processing_costs=[[12 * (i + 1) for i in range(P)], [10 * (i + 1) for i in range(P)]],
                stock_costs=[1 * (i + 1) for i in range(P)])


def multi_product_increasing_costs(num_products=2, **kw):
    """SupplyChainMultiProduct_IncreasingCosts (supplychain_multiproduct_env.py:117-155)."""
    return multi_product(num_products=num_products, **_increasing_costs(num_products), **kw)


def _demand_by_product(P, demand_std, demand_perturb_norm, wrap_std):
    """Per-product demand models of the *_DemConfigByProd classes (:157-209, :211-273): product 0
    seasonal (0,400) 4 peaks avg (100,300); product 1 plain (0,300); product 2 seasonal 2 peaks."""
    assert 1 <= P <= 3
    ranges, stds, peaks, avgs = [(0, 400)], [demand_std], [4], [(100, 300)]
    if P > 1:
        ranges.append((0, 300)); stds.append([demand_std] if wrap_std else demand_std)
        peaks.append(None); avgs.append(None)
    if P > 2:
        ranges.append((0, 400)); stds.append([demand_std] if wrap_std else demand_std)
        peaks.append(2); avgs.append((100, 300))
    return dict(demand_config_by_product=True, demand_range=ranges, demand_std=stds, demand_sen_peaks=peaks,
                avg_demand_range=avgs, demand_perturb_norm=[demand_perturb_norm] * P)


def multi_product_dem_by_prod(num_products=2, demand_std=None, demand_perturb_norm=False, **kw):
    return multi_product(num_products=num_products,
                         **_demand_by_product(num_products, demand_std, demand_perturb_norm, False), **kw)


def multi_product_dem_by_prod_inc_costs(num_products=2, demand_std=None, demand_perturb_norm=False, **kw):
    return multi_product(num_products=num_products, **_increasing_costs(num_products),
                         **_demand_by_product(num_products, demand_std, demand_perturb_norm, True), **kw)


def main_demo_chain(total_time_steps=360):
    """The 2-2-2-2 chain with stochastic lead times built in the reference's __main__ demo
    (supplychain_env.py:816-875); used as the default `nodes_info` of supplychain-v0."""
    stock_cost, cap, ship = 1, 300, 300
    dc = [[2 * stock_cost] * 2]
    names = [["Supplier 1", "Supplier 2"], ["Factory  1", "Factory  2"], ["Wholesal 1", "Wholesal 2"],
             ["Retailer 1", "Retailer 2"]]
    stocks = [[10, 0], [0, 0], [10, 15], [10, 20]]

    def base(k, i):
        return {"initial_stock": stocks[k][i], "stock_capacity": cap, "stock_cost": stock_cost}
    ech = [[dict(base(0, i), supply_capacity=50, supply_cost=5) for i in range(2)],
           [dict(base(1, i), processing_capacity=50, processing_cost=10) for i in range(2)],
           [base(2, i) for i in range(2)], [base(3, i) for i in range(2)]]
    info = _layered(names, ech, [dc] * 3, [[ship] * 2] * 3)
    kw = dict(num_products=1, demand_range=(10, 20), unmet_demand_cost=54, exceeded_stock_capacity_cost=10,
              exceeded_process_capacity_cost=10, exceeded_ship_capacity_cost=10, processing_ratio=3,
              stochastic_leadtimes=True, avg_leadtime=2, max_leadtime=4, total_time_steps=total_time_steps)
    return info, kw
