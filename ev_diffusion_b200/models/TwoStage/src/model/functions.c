/*
 * TwoStage -- scripts for the function-condition / state-list fixture.
 *
 * Authored for this repository with the stock model API only.  Units charge at random
 * while idle; `fire` runs only for idle units whose energy reached 5 and moves them to
 * the active list; active units burn their heat down and `cool` (condition heat < 1.5)
 * sends them back to idle.  Integer state plus one float that only sees exact halving.
 *
 * Layers:  charge -> fire -> burn -> cool
 */
#ifndef _FUNCTIONS_H_
#define _FUNCTIONS_H_

#include "header.h"

__FLAME_GPU_FUNC__ int charge(xmachine_memory_Unit* agent, RNG_rand48* rand48)
{
	if (rnd<CONTINUOUS>(rand48) < 0.5f)
		agent->energy += 2;
	else
		agent->energy += 1;
	return 0;
}

__FLAME_GPU_FUNC__ int fire(xmachine_memory_Unit* agent)
{
	agent->fired++;
	agent->heat = (float)agent->energy;
	agent->energy = 0;
	return 0;
}

__FLAME_GPU_FUNC__ int burn(xmachine_memory_Unit* agent)
{
	agent->heat = agent->heat * 0.5f;
	return 0;
}

__FLAME_GPU_FUNC__ int cool(xmachine_memory_Unit* agent)
{
	agent->heat = 0.0f;
	return 0;
}

#endif
