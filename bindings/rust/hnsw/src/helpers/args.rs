//! hnsw/src/helpers/args.rs: `lim m` from the command line (eval_glove/src/main.rs:19-28 of the reference)
pub fn parse_args_eval() -> Result<(usize, usize), String> {
    let a: Vec<String> = std::env::args().collect();
    if a.len() < 3 { return Err("expected: lim[int] m[int]".to_string()); }
    let lim = a[1].parse::<usize>().map_err(|e| format!("lim: {e}"))?;
    let m = a[2].parse::<usize>().map_err(|e| format!("m: {e}"))?;
    Ok((lim, m))
}
