// floria-pin-harness — runs the UNMODIFIED floria library on an H-PoP fragment file and dumps what the hot path produced,
// in floria's own formats, so that tools/pin_with_floria.sh can diff it against this repository's output:
//   <out>/local_parts/<j>-<l>-<snp_lo>-<ploidy>.haplosets   written by floria itself (graph_processing.rs:289-300) because the
//                                                            logger is at Debug level
//   <out>/final.haplosets                                     process_reads_for_final_parts + get_hapq on the block partitions
// The fragment file fixes the inputs (no BAM parsing / realignment differences): the parity of the reader
// is a separate question from the parity of the scoring / clustering path pinned here.
//
// usage: floria-pin-harness <frags.hpop> <snp_to_genome_pos.txt> <out_dir> <epsilon> <max_ploidy> <block_length>
use floria::file_reader;
use floria::graph_processing;
use floria::types_structs::{Frag, Options};
use std::fs;

fn main() {
    let a: Vec<String> = std::env::args().collect();
    if a.len() < 7 {
        eprintln!("usage: {} <frags.hpop> <snp_to_genome_pos.txt> <out_dir> <epsilon> <max_ploidy> <block_length>", a[0]);
        std::process::exit(2);
    }
    // Debug level makes get_local_hap_blocks write the local_parts dump (graph_processing.rs:289)
    simple_logger::SimpleLogger::new().with_level(log::LevelFilter::Debug).init().unwrap();
    let mut frags: Vec<Frag> = file_reader::get_frags_container(&a[1]).remove("frag_contig").unwrap();
    // src/bin/floria.rs:289-293: sort by Frag::cmp, then counter_id = index
    frags.sort();
    for (i, f) in frags.iter_mut().enumerate() {
        f.counter_id = i;
    }
    let snp_to_genome_pos: Vec<usize> =
        fs::read_to_string(&a[2]).unwrap().split_whitespace().map(|x| x.parse().unwrap()).collect();
    let options = Options {
        epsilon: a[4].parse().unwrap(),
        max_ploidy: a[5].parse().unwrap(),
        block_length: a[6].parse().unwrap(),
        max_number_solns: 10,      // parse_cmd_line.rs:34
        snp_density: 0.0005,       // parse_cmd_line.rs:38
        stopping_heuristic: true,
        ploidy_sensitivity: 2,     // parse_cmd_line.rs:160
        out_dir: a[3].clone(),
        num_threads: 1,
        ..Default::default()
    };
    fs::create_dir_all(&a[3]).unwrap();
    let hap_graph = graph_processing::generate_hap_graph(&frags, &snp_to_genome_pos, a[3].clone(), &options);
    eprintln!("floria-pin-harness: {} block columns phased", hap_graph.len());
}
