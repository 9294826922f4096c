from murmura_b200.cli import app

app()
