"""Records the stdout of the reference's OWN example programs (examples/*.cpp of parsiad/QuantPDE,
unmodified, built by oracle/Makefile `ref_examples` against oracle/shim) for a set of JSON
configurations: tests/golden/tables/<case>.json (the configuration handed to the program) and
<case>.txt (its stdout: the convergence table of Modules/Utilities/Results.hpp:79-168 and the
solution on the print grid).  tests/test_tables_gpu.py runs examples/bin/*_b200 on the same
configurations and compares byte for byte, the timing column apart.

    python tests/golden/make_tables.py        (needs /root/reference; run in the build container)
"""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "examples")
OUT = os.path.join(HERE, "tables")

# case -> (reference program, configuration or None for programs that take none)
CASES = {
    "vanilla_american_put": ("vanilla_options", {"is_american": True, "is_call": False}),      # configs[0]: the default table, k = 0 .. 5,
    "vanilla_european_call": ("vanilla_options", {"maximum_refinement": 5}),
    "vanilla_digital_put": ("vanilla_options", {"is_digital": True, "is_call": False, "maximum_refinement": 3}),
    "vanilla_american_variable_steps": ("vanilla_options", {"is_american": True, "is_call": False,
                                                            "use_variable_timestepping": True, "minimum_refinement": 1,
                                                            "maximum_refinement": 3}),
    "vanilla_american_call_custom_grid": ("vanilla_options", {
        "is_american": True, "is_call": True, "maximum_refinement": 4, "strike_price": 95.0, "asset_price": 105.0,
        "volatility": 0.35, "dividend_rate": 0.02, "interest_rate": 0.03, "time_to_expiry": 0.75,
        "initial_number_of_timesteps": 10, "print_asset_price_step_size": 7.0,
        "initial_grid": [[0.0, 25.0, 50.0, 70.0, 80.0, 90.0, 95.0, 100.0, 105.0, 110.0, 120.0, 140.0, 180.0, 300.0,
                          1000.0, 10000.0]]}),
    "jump_kou": ("jump_diffusion", {"jump_type": "double_exponential", "minimum_refinement": 3, "maximum_refinement": 6}),   # the rows recorded in jump_diffusion.cpp:184-187
    "jump_merton": ("jump_diffusion", {"minimum_refinement": 1, "maximum_refinement": 4}),
    "hjb_short": ("unequal_borrowing_lending_rates", {"maximum_refinement": 3}),
    "hjb_long": ("unequal_borrowing_lending_rates", {"long_position": True, "maximum_refinement": 3,
                                                     "interest_rate_long": 0.06, "volatility": 0.25}),
    "tutorial": ("tutorial", None),
    # examples/experimental/glwb.cpp: the default DAV2004R mortality table (57 years), and short tables of its own
    "glwb_default": ("glwb", {"maximum_refinement": 4}),
    "glwb_short_tables": ("glwb", {"maximum_refinement": 5, "minimum_refinement": 1, "volatility": 0.25,
                                   "initial_number_of_timesteps": 16, "bonus_rate": 0.05, "withdrawal_rate": 0.06,
                                   "mortality_data": [[0.01, 0.02, 0.04, 0.08, 0.15, 0.3, 0.5, 1.0]],
                                   "penalty_data": [[0.05, 0.04, 0.03, 0.02, 0.01, 0.0, 0.0, 0.0]]}),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for case, (prog, conf) in CASES.items():
        args = [os.path.join(REF, prog)]
        if conf is not None:
            path = os.path.join(OUT, case + ".json")
            with open(path, "w") as f:
                json.dump(conf, f, indent=1, sort_keys=True)
                f.write("\n")
            args.append(path)
        out = subprocess.run(args, check=True, capture_output=True, text=True).stdout
        with open(os.path.join(OUT, case + ".txt"), "w") as f:
            f.write(out)
        print(case, "->", len(out.splitlines()), "lines")


if __name__ == "__main__":
    main()
