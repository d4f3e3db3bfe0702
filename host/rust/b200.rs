//! `src/b200.rs` -- the shim a maintainer of the reference adds (with `mod b200;` in `main.rs` and `build.rs` next to
//! `Cargo.toml`) to send `WorldState::lookahead`, `mind::score` and the kickoff family to `libleafline_b200.so`.
//!
//! UNTESTED SOURCE: the build image of this repository has no Rust toolchain.  The same symbols, with the same
//! signatures, are exercised in-tree through `ctypes` (`leafline_b200/_abi.py`) and C++ (`host/leafline.hpp`);
//! `tests/test_abi.py` keeps header, binding and exported symbols in step.  Types and field names are the
//! reference's (`life.rs:15-20, 77-83, 157-176`, `identity.rs:7-49`, `space.rs:1-4, 128`).

use std::ffi::CStr;
use std::os::raw::{c_char, c_int};
use std::ptr;

use identity::{Agent, JobDescription, Team};
use life::{Commit, Patch, WorldState};
use space::{Locale, Pinfield};

const LL_NONE: u8 = 0xFF;
const LL_MAX_VARIATION: usize = 16;

#[repr(C)]
#[derive(Copy, Clone)]
pub struct LlWorld {
    plane: [u64; 12],        // orange servants .. figurehead, blue servants .. figurehead (life.rs:161-173)
    initiative: u8,          // 0 Orange, 1 Blue
    service_eligibility: u8, // life.rs:178-181
    passing_by: u8,          // 0xFF = None, else rank << 4 | file (space.rs:1-4)
    pad: [u8; 5],
}

#[repr(C)]
#[derive(Copy, Clone)]
pub struct LlCommit {
    team: u8,
    job: u8,
    whence: u8,
    whither: u8,
    hospitalization: u8, // 0xFF = None, else team << 3 | job
    ascension: u8,
    pad: [u8; 2],
    tree: LlWorld,
}

#[repr(C)]
#[derive(Copy, Clone)]
pub struct LlPatch {
    team: u8,
    job: u8,
    whence: u8,
    whither: u8,
}

#[repr(C)]
#[derive(Copy, Clone)]
pub struct LlForecast {
    commit: LlCommit,
    score: f32,
    variation_len: u32,
    variation: [LlPatch; LL_MAX_VARIATION],
}

pub enum LlCtx {}
pub enum LlMctx {}              // ll_mctx: several GPUs sharing one tree

extern "C" {
    fn ll_init(device: c_int, ctx: *mut *mut LlCtx) -> c_int;
    fn ll_shutdown(ctx: *mut LlCtx) -> c_int;
    fn ll_last_error() -> *const c_char;
    fn ll_score_batch(ctx: *mut LlCtx, worlds: *const LlWorld, n: usize, scores: *mut f32) -> c_int;
    fn ll_lookahead_batch(ctx: *mut LlCtx, worlds: *const LlWorld, n: usize, nihilistically: c_int, offsets: *mut u32,
                          commits: *mut LlCommit, commits_cap: usize) -> c_int;
    fn ll_in_critical_endangerment_batch(ctx: *mut LlCtx, worlds: *const LlWorld, teams: *const u8, n: usize,
                                         endangered: *mut u8) -> c_int;
    fn ll_perft(ctx: *mut LlCtx, root: *const LlWorld, depth: c_int, shard_rank: c_int, shard_world: c_int, leaves: *mut u64,
                per_root: *mut u64, root_cap: u32, n_roots: *mut u32) -> c_int;
    fn ll_horizon_batch(ctx: *mut LlCtx, frontier: *const LlWorld, n: usize, plies: c_int, quiet_extension: c_int,
                        values: *mut f32) -> c_int;
    fn ll_kickoff(ctx: *mut LlCtx, root: *const LlWorld, depth: c_int, quiet_extension: c_int, nihilistically: c_int,
                  shard_rank: c_int, shard_world: c_int, roots: *mut LlCommit, scores: *mut f32, root_cap: u32,
                  n_roots: *mut u32, leaves_scored: *mut u64) -> c_int;
    fn ll_forecast(ctx: *mut LlCtx, root: *const LlWorld, depth: c_int, quiet_extension: c_int, nihilistically: c_int,
                   forecasts: *mut LlForecast, cap: u32, n: *mut u32, leaves_scored: *mut u64) -> c_int;
    fn ll_alpha_beta_forecast(ctx: *mut LlCtx, root: *const LlWorld, depth: c_int, device_plies: c_int, quiet_extension: c_int,
                              nihilistically: c_int, forecasts: *mut LlForecast, cap: u32, n: *mut u32, stats3: *mut u64) -> c_int;
    fn ll_iterative_deepening_forecast(ctx: *mut LlCtx, root: *const LlWorld, timeout_seconds: f64, nihilistically: c_int,
                                       forecasts: *mut LlForecast, cap: u32, n: *mut u32, depth_reached: *mut u32) -> c_int;
    fn ll_fixed_depth_sequence_forecast(ctx: *mut LlCtx, root: *const LlWorld, depths: *const u8, n_depths: u32,
                                        nihilistically: c_int, forecasts: *mut LlForecast, cap: u32, n: *mut u32) -> c_int;
    fn ll_set_deja_vu_bound(ctx: *mut LlCtx, gib: f64) -> c_int;
    // several GPUs, one tree (replaces the thread-per-root-move fan-out of mind.rs:399-437)
    fn ll_init_multi(devices: *const c_int, n: c_int, group: *mut *mut LlMctx) -> c_int;
    fn ll_shutdown_multi(group: *mut LlMctx) -> c_int;
    fn ll_multi_world(group: *mut LlMctx) -> c_int;
    fn ll_perft_multi(group: *mut LlMctx, root: *const LlWorld, depth: c_int, leaves: *mut u64, per_root: *mut u64,
                      root_cap: u32, n_roots: *mut u32) -> c_int;
    fn ll_kickoff_multi(group: *mut LlMctx, root: *const LlWorld, depth: c_int, quiet_extension: c_int, nihilistically: c_int,
                        roots: *mut LlCommit, scores: *mut f32, root_cap: u32, n_roots: *mut u32, leaves_scored: *mut u64) -> c_int;
    fn ll_alpha_beta_forecast_multi(group: *mut LlMctx, root: *const LlWorld, depth: c_int, device_plies: c_int,
                                    quiet_extension: c_int, nihilistically: c_int, forecasts: *mut LlForecast, cap: u32,
                                    n: *mut u32, stats3: *mut u64) -> c_int;
}

// ---- conversions ---------------------------------------------------------------------------------------

fn team_index(team: Team) -> u8 {
    match team {
        Team::Orange => 0,
        Team::Blue => 1,
    }
}

fn job_index(job: JobDescription) -> u8 {
    match job {
        JobDescription::Servant => 0,
        JobDescription::Pony => 1,
        JobDescription::Scholar => 2,
        JobDescription::Cop => 3,
        JobDescription::Princess => 4,
        JobDescription::Figurehead => 5,
    }
}

fn agent_from(team: u8, job: u8) -> Agent {
    let team = if team == 0 { Team::Orange } else { Team::Blue };
    let job_description = match job {
        0 => JobDescription::Servant,
        1 => JobDescription::Pony,
        2 => JobDescription::Scholar,
        3 => JobDescription::Cop,
        4 => JobDescription::Princess,
        5 => JobDescription::Figurehead,
        _ => moral_panic!("the device named a job description that does not exist"),
    };
    Agent { team, job_description }
}

fn agent_maybe(byte: u8) -> Option<Agent> {
    if byte == LL_NONE { None } else { Some(agent_from(byte >> 3, byte & 7)) }
}

fn locale_byte(locale: Locale) -> u8 {
    (locale.rank() << 4) | locale.file()
}

fn locale_from(byte: u8) -> Locale {
    Locale::new(byte >> 4, byte & 0xF)
}

impl<'a> From<&'a WorldState> for LlWorld {
    fn from(w: &'a WorldState) -> LlWorld {
        LlWorld {
            plane: [w.orange_servants.0, w.orange_ponies.0, w.orange_scholars.0, w.orange_cops.0,
                    w.orange_princesses.0, w.orange_figurehead.0,
                    w.blue_servants.0, w.blue_ponies.0, w.blue_scholars.0, w.blue_cops.0,
                    w.blue_princesses.0, w.blue_figurehead.0],
            initiative: team_index(w.initiative),
            service_eligibility: w.service_eligibility,
            passing_by: w.passing_by_locale.map_or(LL_NONE, locale_byte),
            pad: [0; 5],
        }
    }
}

impl From<LlWorld> for WorldState {
    fn from(w: LlWorld) -> WorldState {
        WorldState {
            initiative: if w.initiative == 0 { Team::Orange } else { Team::Blue },
            orange_servants: Pinfield(w.plane[0]),
            orange_ponies: Pinfield(w.plane[1]),
            orange_scholars: Pinfield(w.plane[2]),
            orange_cops: Pinfield(w.plane[3]),
            orange_princesses: Pinfield(w.plane[4]),
            orange_figurehead: Pinfield(w.plane[5]),
            blue_servants: Pinfield(w.plane[6]),
            blue_ponies: Pinfield(w.plane[7]),
            blue_scholars: Pinfield(w.plane[8]),
            blue_cops: Pinfield(w.plane[9]),
            blue_princesses: Pinfield(w.plane[10]),
            blue_figurehead: Pinfield(w.plane[11]),
            service_eligibility: w.service_eligibility,
            passing_by_locale: if w.passing_by == LL_NONE { None } else { Some(locale_from(w.passing_by)) },
        }
    }
}

impl From<LlCommit> for Commit {
    fn from(c: LlCommit) -> Commit {
        Commit {
            patch: Patch { star: agent_from(c.team, c.job), whence: locale_from(c.whence), whither: locale_from(c.whither) },
            tree: WorldState::from(c.tree),
            hospitalization: agent_maybe(c.hospitalization),
            ascension: agent_maybe(c.ascension),
        }
    }
}

impl From<LlPatch> for Patch {
    fn from(p: LlPatch) -> Patch {
        Patch { star: agent_from(p.team, p.job), whence: locale_from(p.whence), whither: locale_from(p.whither) }
    }
}

// ---- one context per host thread (the ABI's contract; kickoff calls from one thread per root move, mind.rs:405) -----

struct Context(*mut LlCtx);

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ll_shutdown(self.0); }
    }
}

thread_local! {
    static CONTEXT: Context = {
        let mut ctx: *mut LlCtx = ptr::null_mut();
        check(unsafe { ll_init(0, &mut ctx) });
        Context(ctx)
    };
}

fn check(status: c_int) {
    if status != 0 {
        let message = unsafe { CStr::from_ptr(ll_last_error()) }.to_string_lossy().into_owned();
        moral_panic!(format!("the B200 leaf layer reported error {} ({})", status, message))
    }
}

fn with_context<R, F: FnOnce(*mut LlCtx) -> R>(f: F) -> R {
    CONTEXT.with(|context| f(context.0))
}

const NO_QUIET: c_int = -1;

fn quiet_argument(extension: Option<u8>) -> c_int {
    extension.map_or(NO_QUIET, |q| q as c_int)
}

// ---- the call sites of INTEGRATION.md ------------------------------------------------------------------------

fn lookahead_of(world: &WorldState, nihilistically: bool) -> Vec<Commit> {
    let raw = LlWorld::from(world);
    let mut offsets = [0u32; 2];
    let mut commits: Vec<LlCommit> = Vec::with_capacity(512);
    with_context(|ctx| {
        check(unsafe {
            ll_lookahead_batch(ctx, &raw, 1, nihilistically as c_int, offsets.as_mut_ptr(), commits.as_mut_ptr(), 512)
        });
    });
    unsafe { commits.set_len(offsets[1] as usize); }
    commits.into_iter().map(Commit::from).collect()
}

/// `WorldState::lookahead` (life.rs:1078): the legal commits, in the reference's order
pub fn lookahead(world: &WorldState) -> Vec<Commit> {
    lookahead_of(world, false)
}

/// `WorldState::reckless_lookahead` (life.rs:1082)
pub fn reckless_lookahead(world: &WorldState) -> Vec<Commit> {
    lookahead_of(world, true)
}

/// `WorldState::in_critical_endangerment` (life.rs:678)
pub fn in_critical_endangerment(world: &WorldState, team: Team) -> bool {
    let raw = LlWorld::from(world);
    let team = team_index(team);
    let mut endangered = 0u8;
    with_context(|ctx| check(unsafe { ll_in_critical_endangerment_batch(ctx, &raw, &team, 1, &mut endangered) }));
    endangered != 0
}

/// `mind::score` (mind.rs:50) of many worlds at once: the intended use is batch hand-off
pub fn scores(worlds: &[WorldState]) -> Vec<f32> {
    let raw: Vec<LlWorld> = worlds.iter().map(LlWorld::from).collect();
    let mut out = vec![0f32; raw.len()];
    with_context(|ctx| check(unsafe { ll_score_batch(ctx, raw.as_ptr(), raw.len(), out.as_mut_ptr()) }));
    out
}

pub fn score(world: WorldState) -> f32 {
    scores(&[world])[0]
}

/// leaf count of the legal-move tree and its divide table (per root move, `lookahead()` order)
pub fn perft(world: &WorldState, depth: u8) -> (u64, Vec<u64>) {
    let raw = LlWorld::from(world);
    let mut leaves = 0u64;
    let mut per_root = vec![0u64; 1024];
    let mut n_roots = 0u32;
    with_context(|ctx| {
        check(unsafe { ll_perft(ctx, &raw, depth as c_int, 0, 1, &mut leaves, per_root.as_mut_ptr(), 1024, &mut n_roots) })
    });
    per_root.truncate(n_roots as usize);
    (leaves, per_root)
}

/// what `α_β_negamax_search(world, plies, ..)` returns for every frontier node (mind.rs:271-364), exact and TT-free:
/// the host driver hands off whole frontiers instead of recursing
pub fn horizon_values(frontier: &[WorldState], plies: u8, extension: Option<u8>) -> Vec<f32> {
    let raw: Vec<LlWorld> = frontier.iter().map(LlWorld::from).collect();
    let mut values = vec![0f32; raw.len()];
    with_context(|ctx| {
        check(unsafe {
            ll_horizon_batch(ctx, raw.as_ptr(), raw.len(), plies as c_int, quiet_argument(extension), values.as_mut_ptr())
        })
    });
    values
}

fn forecasts_from(raw: Vec<LlForecast>) -> Vec<(Commit, f32, Vec<Patch>)> {
    raw.into_iter()
       .map(|f| {
           let variation = f.variation[..f.variation_len as usize].iter().map(|&p| Patch::from(p)).collect();
           (Commit::from(f.commit), f.score, variation)
       })
       .collect()
}

/// `kickoff::<Variation>` (mind.rs:441-448): forecasts sorted by score, best first; full-width on the device
pub fn kickoff(world: &WorldState, depth: u8, extension: Option<u8>, nihilistically: bool) -> Vec<(Commit, f32, Vec<Patch>)> {
    let raw = LlWorld::from(world);
    let mut forecasts: Vec<LlForecast> = Vec::with_capacity(1024);
    let mut n = 0u32;
    with_context(|ctx| {
        check(unsafe {
            ll_forecast(ctx, &raw, depth as c_int, quiet_argument(extension), nihilistically as c_int, forecasts.as_mut_ptr(),
                        1024, &mut n, ptr::null_mut())
        })
    });
    unsafe { forecasts.set_len(n as usize); }
    forecasts_from(forecasts)
}

/// the same forecasts by the host alpha-beta driver over device frontier batches: for depths whose full-width tree
/// does not fit HBM (8 and beyond from the opening)
pub fn alpha_beta_kickoff(world: &WorldState, depth: u8, device_plies: u8, extension: Option<u8>, nihilistically: bool)
                          -> Vec<(Commit, f32, Vec<Patch>)> {
    let raw = LlWorld::from(world);
    let mut forecasts: Vec<LlForecast> = Vec::with_capacity(1024);
    let mut n = 0u32;
    with_context(|ctx| {
        check(unsafe {
            ll_alpha_beta_forecast(ctx, &raw, depth as c_int, device_plies as c_int, quiet_argument(extension),
                                   nihilistically as c_int, forecasts.as_mut_ptr(), 1024, &mut n, ptr::null_mut())
        })
    });
    unsafe { forecasts.set_len(n as usize); }
    forecasts_from(forecasts)
}

/// `iterative_deepening_kickoff` (mind.rs:451-469) -> (forecasts of the deepest iteration that finished in time, its depth)
pub fn iterative_deepening_kickoff(world: &WorldState, timeout_seconds: f64, nihilistically: bool)
                                   -> (Vec<(Commit, f32, Vec<Patch>)>, u8) {
    let raw = LlWorld::from(world);
    let mut forecasts: Vec<LlForecast> = Vec::with_capacity(1024);
    let (mut n, mut depth) = (0u32, 0u32);
    with_context(|ctx| {
        check(unsafe {
            ll_iterative_deepening_forecast(ctx, &raw, timeout_seconds, nihilistically as c_int, forecasts.as_mut_ptr(), 1024,
                                            &mut n, &mut depth)
        })
    });
    unsafe { forecasts.set_len(n as usize); }
    (forecasts_from(forecasts), depth as u8)
}

/// `fixed_depth_sequence_kickoff` (mind.rs:473-490)
pub fn fixed_depth_sequence_kickoff(world: &WorldState, depth_sequence: Vec<u8>, nihilistically: bool)
                                    -> Vec<(Commit, f32, Vec<Patch>)> {
    let raw = LlWorld::from(world);
    let mut forecasts: Vec<LlForecast> = Vec::with_capacity(1024);
    let mut n = 0u32;
    with_context(|ctx| {
        check(unsafe {
            ll_fixed_depth_sequence_forecast(ctx, &raw, depth_sequence.as_ptr(), depth_sequence.len() as u32,
                                             nihilistically as c_int, forecasts.as_mut_ptr(), 1024, &mut n)
        })
    });
    unsafe { forecasts.set_len(n as usize); }
    forecasts_from(forecasts)
}

/// root move scores only, root moves dealt round-robin over `shard_world` processes (one per GPU): `(commit, score)` in
/// `lookahead()` order, `-inf` for the root moves of other shards; the caller max-gathers and sorts (mind.rs:436)
pub fn sharded_root_scores(world: &WorldState, depth: u8, extension: Option<u8>, nihilistically: bool, shard_rank: u32,
                           shard_world: u32) -> Vec<(Commit, f32)> {
    let raw = LlWorld::from(world);
    let mut roots: Vec<LlCommit> = Vec::with_capacity(1024);
    let mut scores = vec![0f32; 1024];
    let mut n = 0u32;
    with_context(|ctx| {
        check(unsafe {
            ll_kickoff(ctx, &raw, depth as c_int, quiet_argument(extension), nihilistically as c_int, shard_rank as c_int,
                       shard_world as c_int, roots.as_mut_ptr(), scores.as_mut_ptr(), 1024, &mut n, ptr::null_mut())
        })
    });
    unsafe { roots.set_len(n as usize); }
    roots.into_iter().map(Commit::from).zip(scores.into_iter()).collect()
}


// ---- several GPUs, one tree ---------------------------------------------------------------------------------------------
// `potentially_timebound_kickoff` spawns a thread per root move and gathers over an mpsc channel (mind.rs:399-437).  The
// group below is that fan-out on the GPUs of one box: the library owns a context, a host thread and a stream per device, the
// NCCL communicator between them and the reduction of the results on the devices; the caller makes ONE call.

/// One process driving `devices` (ll_init_multi).  Not `Sync`: use from one thread at a time.
pub struct Minds {
    group: *mut LlMctx,
}

impl Minds {
    pub fn new(devices: &[i32]) -> Minds {
        let mut group: *mut LlMctx = ptr::null_mut();
        check(unsafe { ll_init_multi(devices.as_ptr(), devices.len() as c_int, &mut group) });
        Minds { group }
    }

    pub fn world(&self) -> usize {
        unsafe { ll_multi_world(self.group) as usize }
    }

    /// leaf count of the legal-move tree, its ply-3 nodes dealt to the GPUs, counts summed on the devices -> (leaves, divide)
    pub fn perft(&self, world: &WorldState, depth: u8) -> (u64, Vec<u64>) {
        let raw = LlWorld::from(world);
        let mut leaves = 0u64;
        let mut per_root = vec![0u64; 1024];
        let mut n = 0u32;
        check(unsafe { ll_perft_multi(self.group, &raw, depth as c_int, &mut leaves, per_root.as_mut_ptr(), 1024, &mut n) });
        per_root.truncate(n as usize);
        (leaves, per_root)
    }

    /// exact root scores in `lookahead()` order: the full-width search dealt by ply-3 nodes, values folded back on the devices
    pub fn root_scores(&self, world: &WorldState, depth: u8, extension: Option<u8>, nihilistically: bool) -> Vec<(Commit, f32)> {
        let raw = LlWorld::from(world);
        let mut roots: Vec<LlCommit> = Vec::with_capacity(1024);
        let mut scores = vec![0f32; 1024];
        let mut n = 0u32;
        check(unsafe {
            ll_kickoff_multi(self.group, &raw, depth as c_int, quiet_argument(extension), nihilistically as c_int, roots.as_mut_ptr(),
                             scores.as_mut_ptr(), 1024, &mut n, ptr::null_mut())
        });
        unsafe { roots.set_len(n as usize); }
        roots.into_iter().map(Commit::from).zip(scores.into_iter()).collect()
    }

    /// `kickoff::<Variation>` by the host alpha-beta driver, root move i searched on GPU i % world (mind.rs:399-414)
    pub fn kickoff(&self, world: &WorldState, depth: u8, device_plies: u8, extension: Option<u8>, nihilistically: bool)
                   -> Vec<(Commit, f32, Vec<Patch>)> {
        let raw = LlWorld::from(world);
        let mut forecasts: Vec<LlForecast> = Vec::with_capacity(1024);
        let mut n = 0u32;
        check(unsafe {
            ll_alpha_beta_forecast_multi(self.group, &raw, depth as c_int, device_plies as c_int, quiet_argument(extension),
                                         nihilistically as c_int, forecasts.as_mut_ptr(), 1024, &mut n, ptr::null_mut())
        });
        unsafe { forecasts.set_len(n as usize); }
        forecasts_from(forecasts)
    }
}

impl Drop for Minds {
    fn drop(&mut self) {
        unsafe { ll_shutdown_multi(self.group); }
    }
}

/// the `déjà_vu_bound` of the reference's kickoff functions (GiB of memory bank, mind.rs:367-372) for this thread's context
pub fn set_deja_vu_bound(gib: f32) {
    with_context(|ctx| check(unsafe { ll_set_deja_vu_bound(ctx, gib as f64) }));
}
