def get_cpu_info():
    return {"brand_raw": "unknown"}
