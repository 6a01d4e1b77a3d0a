"""Command line front end with the reference's arguments (/root/reference/run_rhe.py:161-201).

    python run_rhe.py --config example.txt          # INI file, section [PyRHE_Config]
    torchrun --nproc-per-node 8 run_rhe.py ...      # one process per GPU, blocks sharded

Flags, defaults, the config-file overlay and the log text are kept; `--device` is accepted but
the block path always runs on CUDA (cuda:LOCAL_RANK or --cuda_num).  Extra: `--kernel_path`.
"""
import argparse
import configparser
import os
import time

from pyrhe.src.models.genie import GENIE, StreamingGENIE
from pyrhe.src.models.rhe import RHE, StreamingRHE
from pyrhe.src.models.rhe_dom import RHE_DOM, StreamingRHE_DOM
from pyrhe.src.util import Logger

MODELS = {("rhe", False): RHE, ("rhe", True): StreamingRHE,
          ("rhe_dom", False): RHE_DOM, ("rhe_dom", True): StreamingRHE_DOM,
          ("genie", False): GENIE, ("genie", True): StreamingGENIE}
BANNER = ["##################################", "#                                #",
          "#          PyRHE (v1.0.0)        #", "#                                #",
          "##################################"]


def build_parser():
    p = argparse.ArgumentParser(description="PyRHE")
    p.add_argument("--model", type=str, default="rhe", choices=["rhe", "genie", "rhe_dom"])
    p.add_argument("--genie_model", type=str, default="G+GxE+NxE", choices=["G", "G+GxE", "G+GxE+NxE"])
    p.add_argument("--streaming", action="store_true", help="use streaming version")
    p.add_argument("--trace", "-tr", action="store_true", help="get the trace estimate")
    p.add_argument("--trace_dir", type=str, default="", help="directory to save the trace information")
    p.add_argument("--benchmark_runtime", action="store_true", help="benchmark the runtime")
    p.add_argument("--genotype", "-g", type=str, help="genotype file path")
    p.add_argument("--phenotype", "-p", type=str, default=None, help="phenotype file path")
    p.add_argument("--covariate", "-c", type=str, default=None, help="Covariate file path")
    p.add_argument("--cov_one_hot_conversion", action="store_true")
    p.add_argument("--categorical_threshhold", type=int, default=100)
    p.add_argument("--env", "-e", type=str, default=None, help="Environment file path")
    p.add_argument("--annotation", "-annot", type=str, default=None, help="Annotation file path")
    p.add_argument("--num_vec", "-k", type=int, default=10, help="The number of random vectors (10 is recommended).")
    p.add_argument("--num_bin", "-b", type=int, default=8, help="Number of bins")
    p.add_argument("--num_workers", type=int, default=8, help="Number of workers")
    p.add_argument("--num_block", "-jn", type=int, default=100, help="The number of jackknife blocks.")
    p.add_argument("--seed", "-s", default=None, help="Random seed")
    p.add_argument("--device", type=str, default="cpu", help="device to use")
    p.add_argument("--cuda_num", type=int, default=None, help="cuda number")
    p.add_argument("--output", "-o", type=str, default="test.out", help="output of the file")
    p.add_argument("--geno_impute_method", type=str, default="binary", choices=["binary", "mean"])
    p.add_argument("--cov_impute_method", type=str, default="ignore", choices=["ignore", "mean"])
    p.add_argument("--samp_prev", default=None)
    p.add_argument("--pop_prev", default=None)
    p.add_argument("--suppress", action="store_true")
    p.add_argument("--debug", action="store_true", help="debug mode")
    p.add_argument("--debug_output", type=str, default="test")
    p.add_argument("--config", type=str, help="Configuration file path")
    p.add_argument("--kernel_path", type=int, default=None, help="0 = SIMT, 1 = tcgen05 (default: library default)")
    return p


def coerce(value, default):
    """run_rhe.py:18-26: config strings take the type of the argparse default."""
    if value.lower() == "none":
        return None
    if isinstance(default, bool):
        return value.lower() in ("true", "1", "yes")
    if isinstance(default, int):
        return int(value)
    return value


def apply_config(args):
    cfg = configparser.ConfigParser()
    cfg.read(args.config)
    section = dict(cfg.items("PyRHE_Config"))
    for key, default in vars(args).items():
        if key in section:
            setattr(args, key, coerce(section[key], default))


def main(args):
    rank0 = int(os.environ.get("RANK", "0")) == 0
    log = Logger(output_file=args.output if rank0 else None, suppress=args.suppress or not rank0, debug_mode=args.debug)
    for line in BANNER:
        log._log(line)
    log._log("\n")
    log._log("Active essential options:")
    for flag, value in (("-g (genotype)", args.genotype), ("-annot (annotation)", args.annotation),
                        ("-p (phenotype)", args.phenotype), ("-c (covariates)", args.covariate),
                        ("-o (output)", args.output), ("-k (# random vectors)", args.num_vec),
                        ("-jn (# jackknife blocks)", args.num_block), ("--num_workers", args.num_workers),
                        ("--device", args.device), ("--geno_impute_method", args.geno_impute_method),
                        ("--cov_impute_method", args.cov_impute_method)):
        log._log(f"\t{flag} {value}")
    log._log("\n")
    log._debug(args)
    if (args.samp_prev is not None) != (args.pop_prev is not None):
        raise ValueError("Must set both or neither of --samp-prev and --pop-prev.")
    if args.annotation is None:
        data_dir = os.environ.get("DATA_DIR", ".")
        args.annotation = f"{data_dir}/annot/annot_{args.num_bin}"
    params = dict(model=args.model, geno_file=args.genotype, annot_file=args.annotation, pheno_file=args.phenotype,
                  cov_file=args.covariate, num_jack=args.num_block, num_bin=args.num_bin, num_random_vec=args.num_vec,
                  geno_impute_method=args.geno_impute_method, cov_impute_method=args.cov_impute_method,
                  cov_one_hot_conversion=args.cov_one_hot_conversion,
                  categorical_threshhold=args.categorical_threshhold, device=args.device, cuda_num=args.cuda_num,
                  multiprocessing=args.num_workers > 1, num_workers=args.num_workers, seed=args.seed,
                  get_trace=args.trace, trace_dir=args.trace_dir, samp_prev=args.samp_prev, pop_prev=args.pop_prev,
                  log=log, kernel_path=args.kernel_path)
    if (args.model, bool(args.streaming)) not in MODELS:
        raise ValueError("Unsupported Model")
    if args.model == "genie":
        params.update(env_file=args.env, genie_model=args.genie_model)
    model = MODELS[(args.model, bool(args.streaming))](**params)
    results, runtime = {}, 0.0
    for trait in range(model.num_traits):
        start = time.time()
        res = model(trait=trait)
        runtime = time.time() - start
        results[f"Trait{trait}"] = {**res, "runtime": runtime}
    log._log("Runtime: ", runtime)
    log._save_log()
    return results


if __name__ == "__main__":
    parser = build_parser()
    args = parser.parse_args()
    if args.config:
        apply_config(args)
    main(args)
