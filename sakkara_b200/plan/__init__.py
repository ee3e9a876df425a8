"""Lowering of built models to flat plans: logical (ModelPlan) and kernel-level (KernelPlan)."""
from sakkara_b200.plan.model_plan import ModelPlan, Block, Term, PriorParam, LikelihoodSpec
from sakkara_b200.plan.lower import lower_model, UnsupportedModelError
