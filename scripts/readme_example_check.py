import re,sys
sys.path.insert(0,'.')
src=open('README.md').read()
exec(re.search(r"```python\n(.*?)```",src,re.S).group(1))
