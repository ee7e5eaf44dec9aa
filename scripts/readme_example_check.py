"""Runs every ```python block of README.md in order, in one namespace (on the GPU box): the README's examples must work as printed."""
import re
import sys

sys.path.insert(0, ".")
ns = {}
for i, block in enumerate(re.findall(r"```python\n(.*?)```", open("README.md").read(), re.S)):
    print(f"--- README python block {i}")
    exec(block, ns)
print("planned torque:", ns.get("u"))
print("README examples OK")
