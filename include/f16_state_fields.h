/* f16_state_fields.h - field order of one environment's packed flight-dynamics state.
 *
 * Used by f16_get_state / f16_set_state (include/f16_b200.h) and by the CPU oracle's
 * f16o_fdm_pack_state, so a test can lift any oracle state into the CUDA library (teacher forcing)
 * and compare the two after one env-step. Every field is exchanged as a double regardless of the
 * precision mode the library stores it in.
 *
 * The fields are exactly what must survive between two FGFDMExec::Run() frames for aircraft/f16
 * (SURVEY.md Appendix A.9): Propagate's state vector and integrator histories, the Accelerations
 * outputs consumed one frame later, the Auxiliary outputs the FCS reads one frame late, the FCS
 * actuator / PID memories and the turbine spool state.
 */
#ifndef F16_STATE_FIELDS_H
#define F16_STATE_FIELDS_H

enum f16_state_field {
  /* FGPropagate VState */
  F16S_Q0 = 0, F16S_Q1, F16S_Q2, F16S_Q3,       /* qAttitudeECI (scalar first) */
  F16S_WI_X, F16S_WI_Y, F16S_WI_Z,              /* vPQRi, rad/s, body axes */
  F16S_RI_X, F16S_RI_Y, F16S_RI_Z,              /* vInertialPosition, ft, ECI */
  F16S_VI_X, F16S_VI_Y, F16S_VI_Z,              /* vInertialVelocity, ft/s, ECI */
  F16S_EPA,                                     /* earth position angle, rad */
  /* integrator histories (deque entries that are still read) */
  F16S_VI1_X, F16S_VI1_Y, F16S_VI1_Z,           /* dqInertialVelocity[0]: v_i one frame back */
  F16S_VI2_X, F16S_VI2_Y, F16S_VI2_Z,           /* dqInertialVelocity[1]: v_i two frames back */
  F16S_AI0_X, F16S_AI0_Y, F16S_AI0_Z,           /* FGAccelerations vUVWidot of the last frame */
  F16S_AI1_X, F16S_AI1_Y, F16S_AI1_Z,           /* dqUVWidot[0]: vUVWidot one frame further back */
  F16S_WDOT_X, F16S_WDOT_Y, F16S_WDOT_Z,        /* FGAccelerations vPQRidot of the last frame */
  F16S_ABODY_X, F16S_ABODY_Y, F16S_ABODY_Z,     /* FGAccelerations vBodyAccel of the last frame */
  /* FGAuxiliary outputs the FCS reads one frame late */
  F16S_PQR_X, F16S_PQR_Y, F16S_PQR_Z,           /* velocities/{p,q,r}-aero-rad_sec */
  F16S_ALPHA,                                   /* aero/alpha-rad */
  F16S_MACH,                                    /* velocities/mach */
  F16S_VC_KTS,                                  /* velocities/vc-kts */
  F16S_VG,                                      /* velocities/vg-fps */
  F16S_NPY, F16S_NPZ,                           /* accelerations/n-pilot-{y,z}-norm */
  /* FGFCS memories */
  F16S_TEF,                                     /* fcs/tef-control (kinematic output) */
  F16S_AIL,                                     /* fcs/left-aileron-pos-norm */
  F16S_ELEV,                                    /* fcs/elevator-pos-norm */
  F16S_SB_DEG,                                  /* fcs/speedbrake-pos-deg */
  F16S_ROLL_INPREV, F16S_ROLL_I,                /* fcs/roll-rate-pid: Input_prev, I_out_total */
  F16S_PITCH_INPREV, F16S_PITCH_I,              /* fcs/g-load-pid */
  F16S_YAW_INPREV, F16S_YAW_I,                  /* fcs/yaw-load-pid */
  /* FGTurbine */
  F16S_N2,                                      /* propulsion/engine/n2, percent */
  F16S_AUG,                                     /* Augmentation flag of the last frame (0/1) */
  F16_NUM_STATE_FIELDS
};

#endif /* F16_STATE_FIELDS_H */
